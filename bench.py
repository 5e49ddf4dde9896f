#!/usr/bin/env python
"""Benchmark of the B200-native UVic ESCM 2.9 ocean tracer step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the hot path (isopyc -> vmixc -> MOBI -> tracer, as `mom` sequences
it, source/mom/mom.F:340-389) over one batch of synthetic fields, followed by the halo
exchange (N > 1) and the time-level rotation.  The default workload is the configuration
BASELINE.json quotes its metric on: the 100x100x19 ocean (imt=102, jmt=102, km=19) with
isopycnal mixing + GM + FCT advection and the full MOBI tracer set of run/mk.in (nt=37,
nsrc=35).  With N GPUs the global grid grows to 100*N interior rows (weak scaling): every rank
owns a 100-row latitude slab of the same shape and exchanges 2-row halos over NCCL.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (a restatement of the
reference Fortran, which cannot be built here -- no Fortran compiler, no netCDF, no input
data) on the host cores, one serial replica per core, on the same workload.
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
    # name: (imt, rows per GPU, km, nt, mobi)
    "uvic100_mobi37": dict(imt=102, rows=100, km=19, nt=37, mobi=1,
                           desc="UVic 2.9 100x100x19, isopycnal mixing + GM + FCT + full MOBI tracer set (run/mk.in), nt=37"),
    "uvic100_ts": dict(imt=102, rows=100, km=19, nt=2, mobi=0, desc="UVic 2.9 100x100x19, T,S only, isopyc + GM + FCT"),
    "one_deg_37": dict(imt=362, rows=180, km=30, nt=37, mobi=1,
                       desc="synthetic 1 deg 360x180x30, full MOBI tracer set (profiling size)"),
    "half_deg_40": dict(imt=722, rows=360, km=40, nt=40, mobi=1,
                        desc="synthetic 0.5 deg 720x360x40, 40 tracers (37 MOBI + 3 passive), isopyc + FCT + invtri"),
    "tenth_deg_slab_40": dict(imt=3602, rows=225, km=60, nt=40, mobi=1,
                              desc="synthetic 0.1 deg 3600x(225 rows per GPU)x60 latitude slab, 40 tracers"),
}


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
    """One GPU: the named grid.  N GPUs (weak scaling): N copies of that grid's interior rows stacked in latitude, one per
    GPU (synthetic.stack_bands) -- every slab is the one-GPU problem, neighbours exchange their 2-row halos."""
    w = WORKLOADS[wl]
    base = pkg.synthetic.make_case(imt=w["imt"], jmt=2 + w["rows"], km=w["km"], nt=w["nt"])
    return pkg.synthetic.stack_bands(base, world, lazy=True)


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
        "k_convect_ts": cells * 32,                    # T,S read + written (worst case)
        "k_convect_tr": cells * 16 * (case.nt - 2),    # worst case: every other tracer read + written
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


def _replica(args):
    wl, nsteps, warm = args
    pkg = load_pkg()
    case = make_case(pkg, wl, 1)
    return oracle_step_time(pkg, case, WORKLOADS[wl]["mobi"], nsteps, warm)


def run_reference(a):
    """--impl reference: the CPU restatement of the reference on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    pkg = load_pkg()
    case = make_case(pkg, a.workload, 1)
    cores = os.cpu_count() or 1
    nrep = max(1, min(cores, 64))
    # each replica times `steps` serial steps of the same workload after `warmup` untimed ones;
    # bounded so the whole arm ends within a few minutes
    nsteps = max(1, min(a.steps, 8))
    warm = max(1, min(a.warmup, 1))
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(nrep) as pool:
        per = pool.map(_replica, [(a.workload, nsteps, warm)] * nrep)
    wall = time.perf_counter() - t0
    tmax = max(per)
    units = units_per_step(case)
    value = nrep * units / tmax / 1e9
    w = WORKLOADS[a.workload]
    line = {
        "impl": "reference", "metric": "tracer cell-updates/sec", "value": value, "unit": "G cell*tracer/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tmax / nrep, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": a.workload, "desc": w["desc"], "grid": [case.imt, case.jmt, case.km], "nt": case.nt},
        "cpu_baseline": {"value": value, "unit": "G cell*tracer/s", "cores": nrep, "kind": "port",
                         "sample": f"{nrep} independent serial replicas (the reference is a serial code) x {nsteps} full steps of the "
                                   f"workload after {warm} warm-up; oracle/ C restatement, gcc -O3 -march=x86-64-v3 (the reference builds with -O3, run/mk.ver); "
                                   f"single-replica step {1e3 * min(per):.0f}-{1e3 * tmax:.0f} ms; wall {wall:.0f} s"},
        "e2e": {"value": value, "unit": "G cell*tracer/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="uvic100_mobi37", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    pkg = load_pkg()
    w = WORKLOADS[a.workload]
    warmup = max(a.warmup, 3)
    case = make_case(pkg, a.workload, world)
    # slabs of equal estimated work (wet cells + a share for land), not of equal row counts: the kernels skip land, and the
    # synthetic geography has polar land caps (UVIC_B200_EQUAL_ROWS=1 restores equal row counts)
    # the stacked weak-scaling grid gives every slab the same rows and the same work; for a real (uneven) geography
    # slab.partition_rows_balanced cuts the rows by wet-cell count instead (UVIC_B200_BALANCED=1)
    if world > 1 and os.environ.get("UVIC_B200_BALANCED") == "1":
        parts = pkg.slab.partition_rows_balanced(case["kmt"], world, case.km, float(os.environ.get("UVIC_B200_LAND_COST", "0.5")))
    else:
        parts = pkg.slab.partition_rows(case.jmt, world)
    jlo, jhi = parts[rank]
    # O_fourfil is on in run/mk.in; the synthetic fine grids skip it (the reference's filter
    # tables are fixed-size, source/common/index.h:34)
    # ... and so does the stacked weak-scaling grid, whose polar rows repeat inside the domain
    fourfil = 1 if (a.workload.startswith("uvic100") and world == 1) else 0
    ctx = pkg.TracerContext(case, jlo=jlo, jhi=jhi, device=local, mobi=w["mobi"], fourfil=fourfil)
    ctx.load_state()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    halo = pkg.slab.HaloExchanger(case.jmt, rank, world, dist, parts=parts)
    tviews = {}

    def tp1_tensor():
        p = ctx.t_ptr(+1)
        if p not in tviews:
            tviews[p] = pkg.slab.device_tensor(p, ctx.shape_t(), local)
        return tviews[p]

    state = {"itt": 0}

    halo_stream = torch.cuda.Stream(device=local) if world > 1 else None
    halo_sync = os.environ.get("UVIC_B200_HALO_SYNC") == "1"

    def one_step():
        state["itt"] += 1
        # the host knows its schedule (mixing step every nmix-th itt, source/mom/mom.F:111-146) and says so: MOBI look-ahead
        ctx.step(leapfrog=pkg.timestep.is_leapfrog(state["itt"], 16), next_leapfrog=pkg.timestep.is_leapfrog(state["itt"] + 1, 16))
        if world > 1 and halo_sync:
            halo.exchange(tp1_tensor())        # A/B switch: the exchange in line on the launch stream
        elif world > 1:
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

    for _ in range(warmup):
        one_step()
    # ---- timed region: device resident (production path: MOBI overlapped on its side stream) ----
    barrier()
    l0 = ctx.kernel_launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(a.steps):
        one_step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches - l0
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
    ms_ranks = [ms / a.steps]
    if world > 1:
        tt = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        allms = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allms, tt)
        ms_ranks = [float(x.item()) / a.steps for x in allms]
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    units = units_per_step(case)
    ms_step = ms / a.steps
    value = units / (ms_step * 1e-3) / 1e9

    # ---- end to end: host buffers through the reference-facing C ABI call ---------------
    e2e = None
    if not a.no_e2e:
        sl = lambda n: np.ascontiguousarray(pkg.api.slab_slice(n, case[n], ctx.jbase, ctx.jl, case))

        def pinned(x):
            t = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = x
            return t

        h_vet, h_vnt, h_vbt = pinned(sl("adv_vet")), pinned(sl("adv_vnt")), pinned(sl("adv_vbt"))
        h_stf, h_btf = pinned(sl("stf")), pinned(sl("btf"))
        h_out = torch.empty(ctx.shape_t(), dtype=torch.float64, pin_memory=True)
        h2d = sum(x.numel() * 8 for x in (h_vet, h_vnt, h_vbt, h_stf, h_btf))
        d2h = h_out.numel() * 8

        def e2e_step():
            state["itt"] += 1
            lf = pkg.timestep.is_leapfrog(state["itt"], 16)
            # t(tau-1), t(tau) stay resident (NULL = keep); velocities and vertical b.c. come from the host,
            # t(tau+1) goes back to the host: what the Fortran shim moves every step
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
        for _ in range(a.steps):
            e2e_step()
        ev1.record(stream)
        barrier()
        ms_e = ev0.elapsed_time(ev1)
        wall_e = (time.perf_counter() - t0) * 1e3
        ms_e = max(ms_e, wall_e)   # the call is synchronous: host wall clock bounds it
        if world > 1:
            tt = torch.tensor([ms_e], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt.item())
        # what the link gives: one pinned D2H / H2D copy of the size the step moves, timed alone
        def link_gbs(dst, src):
            dst.copy_(src, non_blocking=True)   # untimed first touch
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(5):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            return 5 * src.numel() * 8 / (time.perf_counter() - t1) / 1e9

        d_tmp = torch.empty(h_out.shape, dtype=torch.float64, device=f"cuda:{local}")
        pcie = {"d2h_gbs": round(link_gbs(h_out, d_tmp), 1), "h2d_gbs": round(link_gbs(d_tmp, h_out), 1)}
        del d_tmp
        e2e = {"value": units / (ms_e / a.steps * 1e-3) / 1e9, "unit": "G cell*tracer/s", "h2d_bytes_per_step": int(h2d),
               "pcie": pcie, "d2h_floor_ms": round(d2h / (pcie["d2h_gbs"] * 1e9) * 1e3, 3),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / a.steps,
               "note": "uvic_b200_tracer_step: adv velocities + stf/btf H2D from pinned memory and the whole t(tau+1) D2H every step, "
                       "copies on their own streams under the kernels; MOBI of step n+1 runs while t(tau+1) of step n travels"}

        # ---- the same step with setvbc / set_sbc on the device: per step the host sends the advective
        # velocities and, once per ocean segment, the coupler's sbc array; it receives T and S of t(tau+1) every step (the
        # density clinic / loadmw need) and the sbc array with the averaged surface accumulators at the end of a segment
        # (segtim = 5 days, dtts = 1.25 days: ntspos = 4, run/control.in:3-4); the other tracers stay resident
        nsbc = 2 * case.nt + 4
        ctx.sbc_setup(nsbc, np.arange(1, case.nt + 1, dtype=np.int32), np.arange(case.nt + 1, 2 * case.nt + 1, dtype=np.int32))
        h_sbc = pinned(np.zeros((nsbc, ctx.jl, case.imt)))
        h_bhf = pinned(np.zeros((ctx.jl, case.imt)))
        h_sbc_out = torch.empty((nsbc, ctx.jl, case.imt), dtype=torch.float64, pin_memory=True)
        h_ts = torch.empty((2,) + tuple(ctx.shape_t()[1:]), dtype=torch.float64, pin_memory=True)
        ntspos = 4

        def coupled_step():
            state["itt"] += 1
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
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(nc):
            coupled_step()
        ev1.record(stream)
        barrier()
        ms_c = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
        if world > 1:
            tt = torch.tensor([ms_c], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_c = float(tt.item())
        vel_b = sum(x.numel() * 8 for x in (h_vet, h_vnt, h_vbt))
        e2e["coupled"] = {"value": units / (ms_c / nc * 1e-3) / 1e9, "unit": "G cell*tracer/s", "ms_per_step": ms_c / nc, "steps": nc,
                          "h2d_bytes_per_step": int(vel_b + (h_sbc.numel() + h_bhf.numel()) * 8 / ntspos),
                          "d2h_bytes_per_step": int(h_ts.numel() * 8 + h_sbc_out.numel() * 8 / ntspos),
                          "note": "uvic_b200_tracer_step_coupled: setvbc / set_sbc on the device; velocities in and T,S out every "
                                  "step, the sbc array in / out once per 4-step ocean segment; other tracers resident"}

    # ---- the next row of SURVEY 8(f): the baroclinic momentum step (uvic_b200_clinic) on the same grid, device resident,
    # timed on its own (one GPU): rho = state(t(tau)), smf / bmf, U-cell advective velocities, u(tau+1), zu
    clinic = None
    if world == 1 and not a.workload.startswith("tenth") and os.environ.get("UVIC_B200_BENCH_CLINIC", "1") == "1":
        pkg.synthetic.add_momentum(case)
        slc = lambda n: np.ascontiguousarray(pkg.api.slab_slice(n, case[n], ctx.jbase, ctx.jl, case))
        ctx.clinic_setup(case, fourfil=bool(fourfil))
        ctx.upload_u_level(0, slc("u"))
        ctx.upload_u_level(-1, slc("um1"))
        ctx.adv_vel()
        ctx.upload_smf(np.stack([slc("taux"), slc("tauy")]) * slc("umask")[None, :, 0, :])
        c2dtuv = float(case.scalars["c2dtuv"])
        for _ in range(3):
            ctx.clinic(c2dtuv)
        ctx.profile_reset()
        barrier()
        nck = max(a.steps, 20)
        ev0.record(stream)
        for _ in range(nck):
            ctx.clinic(c2dtuv)
        ev1.record(stream)
        barrier()
        ms_ck = ev0.elapsed_time(ev1) / nck
        ctx.profile_enable(True)
        for _ in range(nck):
            ctx.clinic(c2dtuv)
        barrier()
        ctx.profile_enable(False)
        pk = {k: round(1e3 * v[0] / max(v[1], 1), 2) for k, v in ctx.profile().items()
              if k.startswith("k_clinic") or k in ("k_setvbc_mom", "k_state", "k_filuv", "k_filuv_mean")}
        ctx.profile_reset()
        kmu = np.asarray(case["kmu"])[1:-1, 1:-1]
        wet_u = int(kmu.sum())
        all_u = kmu.size * case.km
        # compulsory bytes of the tendency kernel per wet U cell: u(tau) 2 + u(tau-1) 2 + adv_veu/vnu/vbu 3 + visc_ceu,
        # amc_north, amc_south 3 + grad_p 2 reads and du/dt 2 writes (k_clinic_tend); the single marching kernel
        # (k_clinic_column) reads rho instead of grad_p and also carries the masked cells: u(tau-1) 2 reads, u(tau+1) 2 writes
        if "k_clinic_tend" in pk:
            rk, rbytes = "k_clinic_tend", 8 * 14 * wet_u
        else:
            rk, rbytes = "k_clinic_column", 8 * (13 * wet_u + 4 * (all_u - wet_u))
        pkh = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        r_us = pk.get(rk)
        clinic = {"ms_per_step": ms_ck, "steps": nck, "value": 2 * all_u / (ms_ck * 1e-3) / 1e9, "unit": "G U-cell*component/s",
                  "kernels_us": pk,
                  "roofline": {"bound": "hbm", "kernel": rk, "bytes_per_launch": rbytes, "us_per_launch": r_us,
                               "achieved": (rbytes / (r_us * 1e-6) / 1e9) if r_us else None, "peak": pkh, "unit": "GB/s",
                               "frac": (rbytes / (r_us * 1e-6) / 1e9 / pkh) if r_us else None},
                  "note": "uvic_b200_clinic (09/mom/clinic.F with run/mk.in options; filuv as O_fourfil of the workload says), resident fields, CUDA events"}
        if not a.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            os.environ["UVIC_ORACLE_VARIANT"] = "o3"
            import helpers
            oc = helpers.make_oracle(case)
            helpers.oracle_load_momentum(oc, case)
            helpers.oracle_clinic(oc)
            t0 = time.perf_counter()
            for _ in range(3):
                helpers.oracle_clinic(oc)
            cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
            oc.close()
            clinic["cpu_baseline"] = {"ms_per_step": cpu_ms, "value": 2 * all_u / (cpu_ms * 1e-3) / 1e9, "unit": "G U-cell*component/s",
                                      "cores": 1, "kind": "port", "sample": "3 calls of the oracle's adv_vel + state + setvbc + clinic on the same grid"}

    # ---- conservation check on the state the timed steps produced -----------------------
    inv = ctx.inventory(0)

    # ---- load balance: every rank's own step time with the exchange switched off (diagnostic; last, because the
    # halos go stale) -- what the work-balanced partition is tuned against
    solo_ranks = None
    if world > 1:
        barrier()
        ev0.record(stream)
        for _ in range(10):
            state["itt"] += 1
            ctx.step(leapfrog=True, next_leapfrog=True)
            ctx.rotate()
        ev1.record(stream)
        torch.cuda.synchronize()
        tt = torch.tensor([ev0.elapsed_time(ev1) / 10], device=f"cuda:{local}", dtype=torch.float64)
        allms = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allms, tt)
        solo_ranks = [round(float(x.item()), 4) for x in allms]

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    kern = []
    for name, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        per_launch_ms = kms / max(cnt, 1)
        groups = max(1, cnt // a.steps)
        ng = -(-case.nt // groups) if name in PER_TRACER_KERNELS else case.nt
        b = kernel_bytes_per_launch(name, case, ctx, ng)
        kern.append({"kernel": name, "launches": cnt, "ms_total": round(kms, 4), "share": round(kms / tot_ms, 4),
                     "us_per_launch": round(1e3 * per_launch_ms, 2),
                     "gbs": round(b / (per_launch_ms * 1e-3) / 1e9, 1) if b else None})
    top = kern[0]
    top_ng = -(-case.nt // max(1, top["launches"] // a.steps)) if top["kernel"] in PER_TRACER_KERNELS else case.nt
    top_bytes = kernel_bytes_per_launch(top["kernel"], case, ctx, top_ng)
    achieved = top_bytes / (top["us_per_launch"] * 1e-6) / 1e9 if top_bytes else None
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = (json.load(open(tpath)).get(a.workload) or {}).get(top["kernel"])
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": top_bytes, "us_per_launch": top["us_per_launch"],
                "note": "compulsory bytes of the launch as designed / CUDA-event time; traffic = ncu dram bytes of one launch "
                        "(profiles/r01_ncu_traffic.json)"}
    # the same roofline arithmetic for the four largest kernels (the top one runs beside the main stream on a third of
    # the SMs in production; the largest kernel of the main stream is the marching FCT)
    tr_all = {}
    if os.path.exists(tpath):
        tr_all = json.load(open(tpath)).get(a.workload) or {}
    roofline_top = []
    for kq in kern[:4]:
        if kq["gbs"]:
            roofline_top.append({"kernel": kq["kernel"], "bound": "hbm", "achieved": kq["gbs"], "peak": peak, "unit": "GB/s",
                                 "frac": round(kq["gbs"] / peak, 4), "traffic": tr_all.get(kq["kernel"]), "share": kq["share"]})
    # The flux kernels are FP64-issue bound, not HBM bound: attach what the committed `ncu --set full` capture of this
    # workload measured for each kernel (FP64-pipe and issue-slot utilisation, DRAM bytes, registers).  Numbers taken
    # under the profiler, quoted as such; the times above are CUDA events of this run.
    kpath = os.path.join(ROOT, "profiles", "r01_ncu_kernels.json")
    if os.path.exists(kpath):
        nk = json.load(open(kpath)).get(a.workload) or {}
        for kq in kern:
            if kq["kernel"] in nk:
                kq["ncu"] = nk[kq["kernel"]]
        if top["kernel"] in nk:
            roofline["fp64_pipe_pct"] = nk[top["kernel"]].get("fp64_pipe_pct")
            roofline["issue_slot_pct"] = nk[top["kernel"]].get("issue_slot_pct")
    step_bytes = algorithmic_step_bytes(case, w["mobi"]) / world
    step_hbm = {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9, "peak": peak,
                "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak, "unit": "GB/s",
                "note": "SURVEY 8(d) B_step per GPU / step time"}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        # bounded sample of the same workload on one host core (about 10-30 s)
        t_est = 2.5 * units / 7.3e6
        ns = max(1, min(5, int(20.0 / max(t_est, 0.1))))
        sub = case
        sample = f"{ns} full steps of the same workload after 1 warm-up"
        if t_est > 60:
            # large synthetic grids: time a latitude sub-slab and scale
            rows = max(8, int((case.jmt - 2) * 20.0 / t_est))
            sub = pkg.synthetic.make_case(imt=case.imt, jmt=rows + 2, km=case.km, nt=case.nt)
            ns = 1
            sample = f"1 full step of a {rows}-row latitude sub-slab of the workload, throughput per cell"
        tstep = oracle_step_time(pkg, sub, w["mobi"], ns, 1)
        cpu = {"value": units_per_step(sub) / tstep / 1e9, "unit": "G cell*tracer/s", "cores": 1, "kind": "port",
               "sample": sample + "; oracle/ C restatement (gcc -O3 -march=x86-64-v3, the reference's -O3 level), single thread",
               "ms_per_step": 1e3 * tstep}

    line = {
        "metric": "tracer cell-updates/sec", "value": value, "unit": "G cell*tracer/s", "n_gpus": world, "steps": a.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": a.workload, "desc": w["desc"], "grid": [case.imt, case.jmt, case.km], "nt": case.nt,
                   "nsrc": case.nsrc, "rows_per_gpu": w["rows"], "parallelism": f"latitude slabs x{world}, 2-row NCCL halos",
                   "rows_per_slab": [int(b - a + 1) for a, b in parts],
                   "wet_fraction_per_slab": [round(float(np.asarray(case["kmt"])[a - 1:b, 1:-1].sum()) / ((b - a + 1) * (case.imt - 2) * case.km), 3)
                                             for a, b in parts],
                   "l2": "per-step working set (3 time levels + sources + coefficients + FCT scratch) exceeds the 126 MB L2; no explicit flush",
                   "time_stepping": "leapfrog with a forward mixing step every 16th (run/control.in nmix=16)"},
        "sim_years_per_day": 86400.0 / (292.0 * ms_step * 1e-3),
        "roofline": roofline, "roofline_top": roofline_top, "step_hbm": step_hbm, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clk, "ms_per_step_per_rank": [round(x, 4) for x in ms_ranks], "ms_per_step_per_rank_without_exchange": solo_ranks, "ms_per_step_serialised_profile_pass": ms_prof / a.steps, "kernels": kern, "inventory_check": {"finite": bool(np.isfinite(inv).all())},
        "next_rows": {"clinic": clinic},
    }
    _emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
