#!/usr/bin/env python
"""Benchmark of the B200-native UVic ESCM 2.9 ocean tracer step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the hot path (isopyc -> vmixc -> MOBI -> tracer, as `mom` sequences
it, source/mom/mom.F:340-389) over one batch of synthetic fields, followed by the halo
exchange (N > 1) and the time-level rotation.  The default workload is the largest single-GPU
configuration BASELINE.json names: the synthetic 0.5 degree grid 720x360x40 with 40 tracers
(37 MOBI + 3 passive), isopycnal mixing + GM + FCT + invtri (config 4; config 5, the 0.1 degree
grid, needs 8 GPUs: --workload tenth_deg_40).  With N GPUs the SAME global grid is cut into N
latitude slabs through the open ocean (STRONG scaling); neighbours exchange 2-row halos of
t(tau+1) over NCCL, and after the timed region the N-slab result of two steps is compared bit
for bit with the same grid run in one context on rank 0 (`multi_gpu_parity`).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the restatement of the
reference Fortran that tests/test_cpu_refpin.py pins bit for bit to the reference's own code;
the Fortran itself cannot be built here -- no Fortran compiler) on the host cores, one serial
replica per core, on the same workload.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: imt, interior rows of the GLOBAL grid, km, nt, mobi, fourfil (O_fourfil is on in run/mk.in; the synthetic fine
    # grids leave it off: the reference's filter tables are fixed-size, source/common/index.h:34)
    "half_deg_40": dict(imt=722, rows=360, km=40, nt=40, mobi=1, fourfil=0,
                        desc="synthetic 0.5 deg grid 720x360x40, 40 tracers (37 MOBI + 3 passive), isopyc + GM + FCT + invtri (BASELINE config 4)"),
    "uvic100_mobi37": dict(imt=102, rows=100, km=19, nt=37, mobi=1, fourfil=1,
                           desc="UVic 2.9 100x100x19, isopycnal mixing + GM + FCT + full MOBI tracer set with isotopes (run/mk.in), nt=37 (BASELINE config 3)"),
    "uvic100_mobi21": dict(imt=102, rows=100, km=19, nt=21, mobi=1, fourfil=1, options=(),
                           desc="UVic 2.9 100x100x19, isopycnal mixing + GM + FCT + full MOBI tracer set WITHOUT isotopes (O_carbon_13/14, O_mobi_nitrogen_15 off), nt=21 (BASELINE config 2)"),
    "uvic100_ts": dict(imt=102, rows=100, km=19, nt=2, mobi=0, fourfil=1, desc="UVic 2.9 100x100x19, T,S only, isopyc + GM + FCT (BASELINE config 1 physics)"),
    "one_deg_37": dict(imt=362, rows=180, km=30, nt=37, mobi=1, fourfil=0,
                       desc="synthetic 1 deg 360x180x30, full MOBI tracer set (profiling size)"),
    "tenth_deg_40": dict(imt=3602, rows=1800, km=60, nt=40, mobi=1, fourfil=0, bands=8,
                         desc="synthetic 0.1 deg grid 3600x1800x60, 40 MOBI tracers, 8 latitude slabs of 225 rows (BASELINE config 5; needs 8 GPUs)"),
    "tenth_deg_slab_40": dict(imt=3602, rows=225, km=60, nt=40, mobi=1, fourfil=0,
                              desc="one 225-row latitude slab of the synthetic 0.1 deg grid 3600x1800x60, 40 tracers (one GPU's share of config 5)"),
}
DEFAULT_WORKLOAD = "half_deg_40"


_JSON_OUT = None


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_pkg():
    if "uvic29_b200" in sys.modules:
        return sys.modules["uvic29_b200"]
    path = os.path.join(ROOT, "uvic2.9_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("uvic29_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["uvic29_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_case(pkg, wl, world):
    """The named GLOBAL grid, whatever the number of GPUs (strong scaling).  The 0.1 degree grid is 8 bands of 225 rows
    stacked in latitude (synthetic.stack_bands, lazy: a rank holds one 47 GB band on the host, not the 250 GB grid)."""
    w = WORKLOADS[wl]
    bands = w.get("bands", 1)
    if bands > 1:
        if world != bands:
            raise SystemExit(f"bench.py: workload {wl} is {bands} slabs of {w['rows'] // bands} rows: run it with --gpus {bands}")
        base = pkg.synthetic.make_case(imt=w["imt"], jmt=2 + w["rows"] // bands, km=w["km"], nt=w["nt"])
        return pkg.synthetic.stack_bands(base, bands, lazy=True)
    names = pkg.mobi_params.tracer_names_for(options=w["options"]) if "options" in w else None
    return pkg.synthetic.make_case(imt=w["imt"], jmt=2 + w["rows"], km=w["km"], nt=w["nt"], names=names)


def partition(pkg, case, wl, world):
    """Latitude slabs of equal estimated work (wet cells + a share for land: the kernels skip land and the synthetic
    geography has polar land caps); equal row counts for the banded 0.1 degree grid (every band is the same ocean)."""
    if world == 1:
        return [(2, case.jmt - 1)]
    if WORKLOADS[wl].get("bands", 1) > 1 or os.environ.get("UVIC_B200_EQUAL_ROWS") == "1":
        return pkg.slab.partition_rows(case.jmt, world)
    return pkg.slab.partition_rows_balanced(case["kmt"], world, case.km, float(os.environ.get("UVIC_B200_LAND_COST", "0.5")))


def config_dict(a, w, case, world, parts, np):
    """the `config` object: identical for the GPU arm and the reference arm"""
    kmt = np.asarray(case["kmt"])
    return {"workload": a.workload, "desc": w["desc"], "grid": [int(case.imt), int(case.jmt), int(case.km)], "nt": int(case.nt),
            "nsrc": int(case.nsrc), "global_rows": int(case.jmt - 2), "fourfil": int(w["fourfil"]),
            "parallelism": f"latitude slabs x{world}, 2-row halos of t(tau+1) over NCCL" if world > 1 else "one GPU",
            "rows_per_slab": [int(q - p + 1) for p, q in parts],
            "wet_fraction_per_slab": [round(float(kmt[p - 1:q, 1:-1].sum()) / ((q - p + 1) * (case.imt - 2) * case.km), 3) for p, q in parts],
            "l2": "per-step working set (3 time levels + sources + coefficients) exceeds the 126 MB L2; no explicit flush",
            "time_stepping": "leapfrog with a forward mixing step every 16th (run/control.in nmix=16)"}


def units_per_step(case):
    # tracer cell-updates per step: land cells count
    return (case.imt - 2) * (case.jmt - 2) * case.km * case.nt


def algorithmic_step_bytes(case, mobi):
    """B_step = N*[24 nt + 16 nsrc + 8 n_mobi_in + 2*8*C], C = 24 shared fields."""
    n = case.imt * case.km * case.jmt
    nsrc = case.nsrc if mobi else 0
    n_mobi_in = 37 if mobi else 0
    return n * (24 * case.nt + 16 * nsrc + 8 * n_mobi_in + 384)


def kernel_bytes_per_launch(name, case, ctx, ngroup_launch):
    """Compulsory bytes of one launch of each kernel as designed: every
    distinct field it reads once + every field it writes once, for the cells it covers."""
    cells = (case.imt - 2) * case.km * (ctx.jhi - ctx.jlo + 1)
    cells_r = (case.imt - 2) * case.km * (min(case.jmt - 1, ctx.jhi + 1) - max(2, ctx.jlo - 1) + 1)
    g = ngroup_launch
    ocean = float((case["kmt"][ctx.jlo - 1:ctx.jhi, 1:-1]).sum())  # ocean cells in the owned rows
    table = {
        "k_fct_march": cells * (32 * g + 24),          # t(tau-1), t(tau), tendency in; tendency out per tracer; 3 face velocities
        "k_diffuse": cells * (16 * g + 176),           # t(tau-1) in, tendency out per tracer; 22 shared coefficient fields
        "k_fct_tlo": cells_r * (16 * g + 24),          # (two-pass reference variants)
        "k_fct_rfac": cells_r * (72 * g + 24),
        "k_fct_apply": cells * (80 * g + 24),
        "k_update": cells * (80 * g + 176),
        "k_invtri": ocean * (24 * g + 24) + cells * 8 * g,   # wet cells: tendency, t(tau-1), source, a, e, bet in; all cells: t(tau+1) out
        # k_convect_ts / k_convect_tr: data dependent (only the columns that convect are touched): no algorithmic figure
        "k_mobi_column": ocean * 8 * (37 + 15 + 35),   # 37 tracers + 15 pre-pass fields in; 35 sources out
        "k_mobi_ws": ocean * 8 * (37 + 15 + 35),
        "k_mobi_cell": ocean * 8 * (11 + 1 + 15),      # 11 tracers + light in; 15 pre-pass fields out
        "k_mobi_light": ocean * 8 * (4 + 1),
        "k_elements": cells * 8 * (2 + 8),
        "k_isocoef": cells * 8 * (10 + 19),
        "k_gm_faces": cells * 8 * (8 + 2),
        "k_gm_column": cells * 8 * (5 + 4),
        "k_vmix_cbt": cells * 8 * (6 + 1),
        "k_vmix_factor": cells * 8 * (1 + 3),
        "k_gm_total": cells * 8 * (4 + 2),
    }
    return table.get(name)


def kernel_fp64_instr_per_launch(name, case, ctx, ngroup_launch, workload=None):
    """Thread-level FP64 instructions (DFMA + DADD + DMUL + DSETP ...) one launch of a kernel executes, per the committed ncu
    capture (smsp__sass_thread_inst_executed_op_fp64_pred_on.sum, profiles/<tag>_ncu_kernels.json).  MOBI kernels: the count
    per ocean cell of any captured workload, scaled by the ocean cells of this launch.  Other kernels: the count of one launch
    of the SAME workload on one GPU (same launch shape), scaled by this context's share of the rows.  None without a capture."""
    tag = os.environ.get("UVIC_B200_NCU_TAG", "r02")
    kp = os.path.join(ROOT, "profiles", f"{tag}_ncu_kernels.json")
    if not os.path.exists(kp):
        return None
    allk = json.load(open(kp))
    if name.startswith("k_mobi"):
        for wl, ks in allk.items():
            e = ks.get(name) if isinstance(ks, dict) else None
            if e and e.get("fp64_thread_instr_per_ocean_cell"):
                ocean = float((case["kmt"][ctx.jlo - 1:ctx.jhi, 1:-1]).sum())
                return e["fp64_thread_instr_per_ocean_cell"] * ocean
        return None
    e = (allk.get(workload) or {}).get(name)
    if e and e.get("fp64_thread_instr"):
        return e["fp64_thread_instr"] * (ctx.jhi - ctx.jlo + 1) / float(case.jmt - 2)
    return None


PER_TRACER_KERNELS = ("k_fct_march", "k_diffuse", "k_fct_tlo", "k_fct_rfac", "k_fct_apply", "k_update", "k_invtri")

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step_time(pkg, case, mobi, nsteps, warm=1):
    """Seconds per step of the CPU oracle (single thread) on this case; the -O3 build (the reference's run/mk.ver level)."""
    os.environ["UVIC_ORACLE_VARIANT"] = "o3"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_oracle, oracle_rotate, oracle_set_step

    o = make_oracle(case, do_mobi=mobi)
    itt = 0
    times = []
    for s in range(warm + nsteps):
        itt += 1
        lf = pkg.timestep.is_leapfrog(itt, 16)
        oracle_set_step(o, case, lf)
        t0 = time.perf_counter()
        o.call("ora_step")
        dt = time.perf_counter() - t0
        oracle_rotate(o)
        if s >= warm:
            times.append(dt)
    o.close()
    return sum(times) / len(times)


CPU_RATE = 9.0e6     # cell*tracer/s of the oracle on one host core (measured, round 1): only sizes the bounded samples


def cpu_sample(pkg, wl, budget_s):
    """The bounded CPU sample of a workload: the whole grid when one step fits `budget_s` on one core, otherwise a latitude
    sub-slab of the same grid (same imt, km, nt, same synthetic generator), throughput per cell."""
    w = WORKLOADS[wl]
    rows = w["rows"] // w.get("bands", 1)
    t_est = (w["imt"] - 2) * rows * w["km"] * w["nt"] / CPU_RATE
    rows_s = rows if t_est <= budget_s else max(8, int(rows * budget_s / t_est))
    names = pkg.mobi_params.tracer_names_for(options=w["options"]) if "options" in w else None
    case = pkg.synthetic.make_case(imt=w["imt"], jmt=2 + rows_s, km=w["km"], nt=w["nt"], names=names)
    what = "the whole grid" if rows_s == rows else f"a {rows_s}-row latitude sub-slab of the grid (same imt, km, nt), throughput per cell"
    return case, what


def _replica(args):
    wl, nsteps, warm, budget = args
    pkg = load_pkg()
    case, _ = cpu_sample(pkg, wl, budget)
    return oracle_step_time(pkg, case, WORKLOADS[wl]["mobi"], nsteps, warm), units_per_step(case)


def run_reference(a):
    """--impl reference: the CPU restatement of the reference on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    import numpy as np

    pkg = load_pkg()
    w = WORKLOADS[a.workload]
    world = max(1, a.gpus)
    cores = os.cpu_count() or 1
    nrep = max(1, min(cores, 64))
    # every replica times `nsteps` serial steps of its bounded sample after `warm` untimed ones; sized so the whole arm
    # ends within a few minutes whatever --steps says (the line reports what actually ran)
    budget = 12.0
    nsteps = max(1, min(a.steps, 2))
    warm = 1
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(nrep) as pool:
        per = pool.map(_replica, [(a.workload, nsteps, warm, budget)] * nrep)
    wall = time.perf_counter() - t0
    tmax = max(p[0] for p in per)
    units_s = per[0][1]
    value = nrep * units_s / tmax / 1e9
    # the config object of the GPU arm (same keys, same values): needs the global bathymetry for the slab table
    if w.get("bands", 1) > 1:
        gcase = pkg.synthetic.stack_bands(pkg.synthetic.make_case(imt=w["imt"], jmt=2 + w["rows"] // w["bands"], km=w["km"], nt=2), w["bands"], lazy=True)
        gcase.nt, gcase.nsrc = w["nt"], 35 if w["mobi"] else 0
    else:
        gcase = pkg.synthetic.make_case(imt=w["imt"], jmt=2 + w["rows"], km=w["km"], nt=2)
        gcase.nt, gcase.nsrc = w["nt"], (35 if w["mobi"] else 0)
    parts = partition(pkg, gcase, a.workload, world)
    rows_s = (units_s // ((w["imt"] - 2) * w["km"] * w["nt"]))
    sample = (f"{nrep} independent serial replicas (the reference is a serial code), each {warm} warm-up + {nsteps} timed full steps of "
              f"{'the whole grid' if rows_s == w['rows'] else f'a {rows_s}-row latitude sub-slab of the grid (same imt, km, nt; throughput per cell)'}; "
              f"oracle/ C restatement (bitwise equal to the translated reference, tests/test_cpu_refpin.py), gcc -O3 -march=x86-64-v3 "
              f"(the reference builds with -O3, run/mk.ver); single-replica step {1e3 * min(p[0] for p in per):.0f}-{1e3 * tmax:.0f} ms; wall {wall:.0f} s")
    # Is the port a fair stand-in for the reference's own code?  Where the mechanically translated reference travelled with
    # the snapshot (oracle/_ref/libref_s.so), time one `tracer` step of it and of the port on its 34x26x8 grid, same flags.
    check = None
    try:
        import subprocess

        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "time_ref_vs_oracle.py"), "--json"], capture_output=True, text=True, timeout=240)
        if r.returncode == 0 and r.stdout.strip():
            check = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:   # the check is evidence, not the measurement
        check = {"unavailable": str(e)[:200]}
    line = {
        "impl": "reference", "metric": "tracer cell-updates/sec", "value": value, "unit": "G cell*tracer/s", "n_gpus": a.gpus,
        "steps": nsteps, "warmup": warm, "steps_requested": a.steps, "warmup_requested": a.warmup,
        "ms_per_step": 1e3 * units_per_step(gcase) / (value * 1e9), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(a, w, gcase, world, parts, np),
        "cpu_baseline": {"value": value, "unit": "G cell*tracer/s", "cores": nrep, "kind": "port", "sample": sample,
                         "port_vs_translated_reference": check},
        "e2e": {"value": value, "unit": "G cell*tracer/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-slab vs one-context bitwise check (N > 1)")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="the K-step timed region is repeated until this much time is covered")
    a = ap.parse_args()
    # Keep stdout clean for the ONE JSON line: libraries (NCCL prints its version banner there)
    # get stderr as their fd 1; the JSON goes to the saved descriptor.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        return run_reference(a)

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the tracer step has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1 and os.environ.get("UVIC_B200_BENCH_AFFINITY", "1") == "1":
        # one slice of the host cores per rank, in device order: the pinned host buffers of the host-buffer legs are then
        # first-touched on the socket next to the rank's GPU instead of all on one NUMA node
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    pkg = load_pkg()
    w = WORKLOADS[a.workload]
    warmup = max(a.warmup, 3)
    case = make_case(pkg, a.workload, world)
    parts = partition(pkg, case, a.workload, world)
    jlo, jhi = parts[rank]
    fourfil = w["fourfil"]          # the same at every N: the filter is local to a latitude row
    ctx = pkg.TracerContext(case, jlo=jlo, jhi=jhi, device=local, mobi=w["mobi"], fourfil=fourfil)
    ctx.load_state()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    halo = pkg.slab.HaloExchanger(case.jmt, rank, world, dist, parts=parts)
    tviews = {}

    def tp1_tensor(c=None):
        c = c or ctx
        p = c.t_ptr(+1)
        if p not in tviews:
            tviews[p] = pkg.slab.device_tensor(p, c.shape_t(), local)
        return tviews[p]

    state = {"itt": 0}
    relyr0 = float(case.scalars["relyr"])
    dt_years = float(case.scalars["dtts"]) / (365.0 * 86400.0)

    def advance_time():
        # the model year advances every step, as `mom` does (relyr feeds the MOBI declination and the month of the iron
        # deposition); the look-ahead hint carries the NEXT step's value, so MOBI one step ahead is computed for the right time
        state["itt"] += 1
        ctx.set_time(relyr0 + state["itt"] * dt_years, relyr0 + (state["itt"] + 1) * dt_years)

    halo_stream = torch.cuda.Stream(device=local) if world > 1 else None

    def one_step():
        advance_time()
        # the host knows its schedule (mixing step every nmix-th itt, source/mom/mom.F:111-146) and says so: MOBI look-ahead
        ctx.step(leapfrog=pkg.timestep.is_leapfrog(state["itt"], 16), next_leapfrog=pkg.timestep.is_leapfrog(state["itt"] + 1, 16))
        if world > 1:
            # the exchange of t(tau+1) runs on a side stream beside the next step's coefficient / diffusion kernels; the
            # library waits for it before its first advection kernel (the first reader of the new halo rows)
            ev = halo.exchange_async(tp1_tensor(), stream, halo_stream)
            state["halo_ev"] = ev          # keep the event alive until the library has waited for it
            ctx.wait_before_advection(ev.cuda_event)
        ctx.rotate()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([x], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(warmup - 2, 1)):
        one_step()
    barrier()
    ev0.record(stream)
    for _ in range(2):                 # the last two warm-up steps also size the number of timed regions
        one_step()
    ev1.record(stream)
    barrier()
    est_ms = allmax(ev0.elapsed_time(ev1) / 2)
    regions = int(min(200, max(1, -(-a.min_seconds * 1e3 // max(est_ms * a.steps, 1e-3)))))
    # ---- timed: `regions` back-to-back regions of EXACTLY K steps each (device resident, production path: MOBI one step
    # ahead on its side stream), every region bracketed by barrier + synchronize; the value is the mean region ----
    l0 = ctx.kernel_launches
    la0 = ctx.lookahead_stats()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    region_ms = []
    for _ in range(regions):
        barrier()
        ev0.record(stream)
        for _ in range(a.steps):
            one_step()
        ctx.join_side_streams()        # the look-ahead MOBI of the last step belongs to the region
        ev1.record(stream)
        barrier()
        region_ms.append(allmax(ev0.elapsed_time(ev1)))
    own_ms = ev0.elapsed_time(ev1)
    launches = (ctx.kernel_launches - l0) // regions
    la1 = ctx.lookahead_stats()
    ms = sum(region_ms) / len(region_ms)
    # ---- the same K steps again with per-kernel CUDA events (library hooks).  In this pass the
    # library runs MOBI in line on the launch stream, so that no two kernels share the SMs and
    # every event interval is the duration of exactly one kernel; its total is NOT the bench value.
    ctx.profile_reset()
    ctx.profile_enable(True)
    barrier()
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(stream)
    for _ in range(a.steps):
        one_step()
    evp1.record(stream)
    barrier()
    ms_prof = evp0.elapsed_time(evp1)
    clk = clocks.stop() if rank == 0 else None
    ctx.profile_enable(False)
    prof = ctx.profile()
    ms_ranks = [own_ms / a.steps]
    if world > 1:
        tt = torch.tensor([own_ms], device=f"cuda:{local}", dtype=torch.float64)
        allms = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allms, tt)
        ms_ranks = [float(x.item()) / a.steps for x in allms]
    units = units_per_step(case)
    ms_step = ms / a.steps
    value = units / (ms_step * 1e-3) / 1e9

    # ---- end to end: host buffers through the reference-facing C ABI calls ---------------
    e2e = None
    if not a.no_e2e:
        sl = lambda n: np.ascontiguousarray(pkg.api.slab_slice(n, case[n], ctx.jbase, ctx.jl, case))

        def pinned(x):
            t = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = x
            return t

        h_vet, h_vnt, h_vbt = pinned(sl("adv_vet")), pinned(sl("adv_vnt")), pinned(sl("adv_vbt"))
        vel_b = sum(x.numel() * 8 for x in (h_vet, h_vnt, h_vbt))
        # ---- the coupled ocean step (uvic_b200_tracer_step_coupled): setvbc / set_sbc on the device; per step the host sends
        # the advective velocities and, once per ocean segment, the coupler's sbc array; it receives T and S of t(tau+1) every
        # step (the density clinic / loadmw need) and the sbc array with the averaged surface accumulators at the end of a
        # segment (segtim = 5 days, dtts = 1.25 days: ntspos = 4, run/control.in:3-4); the other tracers stay resident between
        # output times -- one whole-state download (what a restart / snapshot needs) is charged to every timed region
        nsbc = 2 * case.nt + 4
        ctx.sbc_setup(nsbc, np.arange(1, case.nt + 1, dtype=np.int32), np.arange(case.nt + 1, 2 * case.nt + 1, dtype=np.int32))
        h_sbc = pinned(np.zeros((nsbc, ctx.jl, case.imt)))
        h_bhf = pinned(np.zeros((ctx.jl, case.imt)))
        h_sbc_out = torch.empty((nsbc, ctx.jl, case.imt), dtype=torch.float64, pin_memory=True)
        h_ts = torch.empty((2,) + tuple(ctx.shape_t()[1:]), dtype=torch.float64, pin_memory=True)
        h_out = torch.empty(ctx.shape_t(), dtype=torch.float64, pin_memory=True)
        ntspos = 4

        def coupled_step():
            advance_time()
            lf = pkg.timestep.is_leapfrog(state["itt"], 16)
            p = (state["itt"] - 1) % ntspos
            ctx.tracer_step_coupled(h_vet.numpy(), h_vnt.numpy(), h_vbt.numpy(), h_sbc.numpy() if p == 0 else None,
                                    h_bhf.numpy() if p == 0 else None, True, p == 0, p == ntspos - 1, ntspos, h_ts.numpy(),
                                    h_sbc_out.numpy(), leapfrog=lf, next_leapfrog=pkg.timestep.is_leapfrog(state["itt"] + 1, 16))
            if world > 1:
                halo.exchange(tp1_tensor())
            ctx.rotate()

        nc = ntspos * max(1, a.steps // ntspos)      # whole segments
        state["itt"] = 0
        for _ in range(ntspos):
            coupled_step()
        reg_e = int(min(20, max(1, -(-a.min_seconds * 1e3 // max(2.0 * est_ms * nc, 1e-3)))))
        tot_ms = 0.0
        for _ in range(reg_e):
            barrier()
            t0 = time.perf_counter()
            ev0.record(stream)
            for _ in range(nc):
                coupled_step()
            ev1.record(stream)
            barrier()
            tot_ms += allmax(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))   # synchronous calls: wall clock bounds it
        # the output step: the whole state (all nt tracers) back to the host.  run/control.in writes full fields every
        # timavgint = 3650 days (time averages, accumulated on the device) and restint = 36500 days: once per 2920 ocean steps
        # at dtts = 1.25 days.  Timed on its own (all ranks at once) and charged to every step at that rate.
        OUTPUT_EVERY = 2920
        barrier()
        t0 = time.perf_counter()
        ctx.download_t(0, out=h_out.numpy())
        barrier()
        ms_full = allmax((time.perf_counter() - t0) * 1e3)
        ms_c = tot_ms / reg_e + nc * ms_full / OUTPUT_EVERY
        h2d_c = int(vel_b + (h_sbc.numel() + h_bhf.numel()) * 8 / ntspos)
        d2h_c = int(h_ts.numel() * 8 + h_sbc_out.numel() * 8 / ntspos + h_out.numel() * 8 / OUTPUT_EVERY)

        # what the link gives: one pinned D2H / H2D copy, timed alone
        def link_gbs(dst, src):
            dst.copy_(src, non_blocking=True)   # untimed first touch
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(5):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            return 5 * src.numel() * 8 / (time.perf_counter() - t1) / 1e9

        d_tmp = torch.empty(h_ts.shape, dtype=torch.float64, device=f"cuda:{local}")
        pcie = {"d2h_gbs": round(link_gbs(h_ts, d_tmp), 1), "h2d_gbs": round(link_gbs(d_tmp, h_ts), 1)}
        del d_tmp
        e2e = {"value": units / (ms_c / nc * 1e-3) / 1e9, "unit": "G cell*tracer/s", "h2d_bytes_per_step": h2d_c * world,
               "d2h_bytes_per_step": d2h_c * world, "ms_per_step": ms_c / nc, "steps": nc, "timed_regions": reg_e, "pcie_one_rank": pcie,
               "api": "uvic_b200_tracer_step_coupled",
               "note": "the coupled ocean step a `mom` host makes: pinned host buffers; every step the advective velocities go in "
                       "and T,S of t(tau+1) come out, the coupler's sbc array in / out once per 4-step ocean segment; the whole-state "
                       "download of an output step (all nt tracers) is timed separately and charged at run/control.in's rate of one "
                       "per 2920 steps; bytes are whole-job",
               "output_step": {"ms": ms_full, "every_steps": OUTPUT_EVERY, "bytes": int(h_out.numel() * 8) * world}}

        # ---- beside it: the full-field variant (uvic_b200_tracer_step): the whole t(tau+1) of all nt tracers back to the host
        # EVERY step -- what a host that keeps the tracers itself (putmw each step) would move
        if os.environ.get("UVIC_B200_BENCH_FULLFIELD", "1") == "1":
            h_stf, h_btf = pinned(sl("stf")), pinned(sl("btf"))
            nf = max(2, min(a.steps, 8))

            def e2e_step():
                advance_time()
                lf = pkg.timestep.is_leapfrog(state["itt"], 16)
                ctx.tracer_step_host(None, None, h_vet.numpy(), h_vnt.numpy(), h_vbt.numpy(), h_stf.numpy(), h_btf.numpy(),
                                     h_out.numpy(), leapfrog=lf, next_leapfrog=pkg.timestep.is_leapfrog(state["itt"] + 1, 16))
                if world > 1:
                    halo.exchange(tp1_tensor())
                ctx.rotate()

            for _ in range(2):
                e2e_step()
            barrier()
            t0 = time.perf_counter()
            ev0.record(stream)
            for _ in range(nf):
                e2e_step()
            ev1.record(stream)
            barrier()
            ms_e = allmax(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
            e2e["full_field_every_step"] = {
                "value": units / (ms_e / nf * 1e-3) / 1e9, "unit": "G cell*tracer/s", "ms_per_step": ms_e / nf, "steps": nf,
                "h2d_bytes_per_step": int(vel_b + (h_stf.numel() + h_btf.numel()) * 8) * world, "d2h_bytes_per_step": int(h_out.numel() * 8) * world,
                "api": "uvic_b200_tracer_step"}

    # ---- N slabs against ONE context, bit for bit (N > 1): both start again from the initial state and take two steps
    parity = None
    full_bytes = 3.3 * case.imt * case.jmt * case.km * case.nt * 8 * 1.6
    if world > 1 and not a.no_parity and not getattr(case, "lazy", None) and full_bytes < 110e9:
        # the same kernels on both sides: the MOBI column kernel is chosen by column count (the warp-specialised one for a
        # slab, the one-thread-per-column one for the whole grid), and with FMA contraction the two agree to 1e-13, not to
        # the bit -- the check is about the decomposition and the NCCL exchange, so it pins the variant
        os.environ["UVIC_B200_MOBI_WS"] = "0"
        ctx.load_state()
        ctx.invalidate_lookahead()
        state["itt"] = 0
        for _ in range(2):
            advance_time()
            ctx.step(leapfrog=True)
            halo.exchange(tp1_tensor())
            ctx.rotate()
        torch.cuda.synchronize()
        mine = pkg.slab.device_tensor(ctx.t_ptr(0), ctx.shape_t(), local)      # after the rotation t(tau) is the new field
        full = torch.empty((case.nt, case.jmt, case.km, case.imt), dtype=torch.float64, device=f"cuda:{local}")
        if rank == 0:
            one = pkg.TracerContext(case, device=local, mobi=w["mobi"], fourfil=fourfil)
            one.load_state()
            one.set_stream(stream.cuda_stream)
            for s_ in range(2):
                one.set_time(relyr0 + (s_ + 1) * dt_years, relyr0 + (s_ + 2) * dt_years)
                one.step(leapfrog=True)
                one.rotate()
            torch.cuda.synchronize()
            full.copy_(pkg.slab.device_tensor(one.t_ptr(0), one.shape_t(), local))
            one.close()
        dist.broadcast(full, src=0)
        own = slice(jlo - ctx.jbase, jhi - ctx.jbase + 1)
        same = bool(torch.equal(mine[:, own], full[:, jlo - 1:jhi]))
        maxdiff = float((mine[:, own] - full[:, jlo - 1:jhi]).abs().max().item())
        flag = torch.tensor([1.0 if same else 0.0, maxdiff], device=f"cuda:{local}", dtype=torch.float64)
        allf = [torch.zeros_like(flag) for _ in range(world)]
        dist.all_gather(allf, flag)
        parity = {"bitwise": all(float(x[0].item()) == 1.0 for x in allf), "max_abs_diff": max(float(x[1].item()) for x in allf),
                  "rows_checked": int(case.jmt - 2), "tracers": int(case.nt), "steps": 2,
                  "how": "every rank compares its owned rows of t after two leapfrog steps (NCCL halo exchange) with the same grid run in ONE "
                         "context on rank 0; MOBI column kernel pinned to the one-thread-per-column variant on both sides"}
        os.environ.pop("UVIC_B200_MOBI_WS", None)
        del full

    # ---- global tracer inventories: per-slab partial sums combined in rank order (fixed order) -------------
    inv_local = ctx.inventory(0)
    inv = pkg.slab.combine_inventories(inv_local, dist, device=f"cuda:{local}") if world > 1 else inv_local

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines --------------------------------------------------------------------------
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp64 = pkg.api.measure_fp64_peak(local)
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    kern = []
    ncu_tag = os.environ.get("UVIC_B200_NCU_TAG", "r02")
    tpath = os.path.join(ROOT, "profiles", f"{ncu_tag}_ncu_traffic.json")
    kpath = os.path.join(ROOT, "profiles", f"{ncu_tag}_ncu_kernels.json")
    tr_all = (json.load(open(tpath)).get(a.workload) or {}) if os.path.exists(tpath) else {}
    nk = (json.load(open(kpath)).get(a.workload) or {}) if os.path.exists(kpath) else {}
    for name, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        per_launch_ms = kms / max(cnt, 1)
        groups = max(1, cnt // a.steps)
        ng = -(-case.nt // groups) if name in PER_TRACER_KERNELS else case.nt
        b = kernel_bytes_per_launch(name, case, ctx, ng)
        row = {"kernel": name, "launches": cnt, "ms_total": round(kms, 4), "share": round(kms / tot_ms, 4),
               "us_per_launch": round(1e3 * per_launch_ms, 2), "bytes_per_launch": b,
               "gbs": round(b / (per_launch_ms * 1e-3) / 1e9, 1) if b else None,
               "hbm_frac": round(b / (per_launch_ms * 1e-3) / 1e9 / peak, 4) if b else None}
        fi = kernel_fp64_instr_per_launch(name, case, ctx, ng, a.workload)
        if fi:
            row["fp64_instr_per_launch"] = fi
            row["fp64_frac"] = round(fi / (per_launch_ms * 1e-3) / fp64["dfma_per_s"], 4)
        if name in nk:
            row["ncu"] = nk[name]
        kern.append(row)
    top = kern[0]

    def roof(kq):
        """one kernel against the roof that bounds it: HBM for the stencil / solve kernels, the measured FP64 issue rate for
        MOBI (arithmetic intensity 13 flop/B, above the machine balance)"""
        # the binding resource is the one the kernel uses the larger fraction of
        if kq.get("fp64_instr_per_launch") and (kq.get("fp64_frac") or 0.0) > (kq.get("hbm_frac") or 0.0):
            ach = kq["fp64_instr_per_launch"] / (kq["us_per_launch"] * 1e-6) / 1e12
            return {"bound": "fp64", "kernel": kq["kernel"], "achieved": ach, "peak": fp64["dfma_per_s"] / 1e12, "unit": "T FP64 instr/s",
                    "frac": ach / (fp64["dfma_per_s"] / 1e12), "traffic": tr_all.get(kq["kernel"]), "share": kq["share"],
                    "us_per_launch": kq["us_per_launch"], "hbm_frac": kq["hbm_frac"], "hbm_gbs": kq["gbs"], "bytes_per_launch": kq["bytes_per_launch"],
                    "peak_source": "measured in this run: uvic_b200_measure_fp64_peak (dependent DFMA chains, thread-level instr/s); "
                                   "instruction count of the launch from the committed ncu capture"}
        if not kq["gbs"]:
            return None
        return {"bound": "hbm", "kernel": kq["kernel"], "achieved": kq["gbs"], "peak": peak, "unit": "GB/s", "frac": round(kq["gbs"] / peak, 4),
                "traffic": tr_all.get(kq["kernel"]), "share": kq["share"], "us_per_launch": kq["us_per_launch"],
                "bytes_per_launch": kq["bytes_per_launch"], "peak_source": peak_src}

    # the dominant kernel that has a roofline figure (the data-dependent convection kernels have none)
    roofline = next((r for r in (roof(kq) for kq in kern) if r), None) or {"bound": "hbm", "kernel": top["kernel"], "achieved": None, "peak": peak,
                                                                             "unit": "GB/s", "frac": None, "traffic": None}
    roofline["note"] = ("dominant kernel of the per-kernel profile pass; achieved = algorithmic bytes (or FP64 instructions) of one launch / its "
                        f"CUDA-event time; traffic = ncu dram bytes of one launch (profiles/{ncu_tag}_ncu_traffic.json)")
    roofline_top = [r for r in (roof(kq) for kq in kern[:6]) if r]
    step_bytes = algorithmic_step_bytes(case, w["mobi"]) / world
    step_hbm = {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9, "peak": peak,
                "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak, "unit": "GB/s",
                "note": "SURVEY 8(d) B_step per GPU / step time"}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        # bounded sample of the same workload on one host core (about 10-30 s)
        sub, what = cpu_sample(pkg, a.workload, 12.0)
        t_est = units_per_step(sub) / CPU_RATE
        ns = max(1, min(5, int(20.0 / max(t_est, 0.1))))
        tstep = oracle_step_time(pkg, sub, w["mobi"], ns, 1)
        cpu = {"value": units_per_step(sub) / tstep / 1e9, "unit": "G cell*tracer/s", "cores": 1, "kind": "port",
               "sample": f"{ns} full steps after 1 warm-up of {what}; oracle/ C restatement (bitwise equal to the translated reference, "
                         "tests/test_cpu_refpin.py), gcc -O3 -march=x86-64-v3 (the reference's -O3 level), single thread",
               "ms_per_step": 1e3 * tstep}

    line = {
        "metric": "tracer cell-updates/sec", "value": value, "unit": "G cell*tracer/s", "n_gpus": world, "steps": a.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(a, w, case, world, parts, np),
        "timed_regions": regions, "timed_seconds_total": sum(region_ms) * 1e-3, "region_ms": [round(x, 3) for x in region_ms[:50]],
        "sim_years_per_day": 86400.0 / (292.0 * ms_step * 1e-3),
        "roofline": roofline, "roofline_top": roofline_top, "step_hbm": step_hbm, "fp64_peak": fp64, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "mobi_lookahead": {"hits": la1[0] - la0[0], "misses": la1[1] - la0[1]},
        "multi_gpu_parity": parity,
        "clocks": clk, "ms_per_step_per_rank": [round(x, 4) for x in ms_ranks], "ms_per_step_serialised_profile_pass": ms_prof / a.steps,
        "kernels": kern, "inventory_check": {"finite": bool(np.isfinite(inv).all()), "global_inventory_first3": [float(x) for x in inv[:3]]},
    }
    _emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
