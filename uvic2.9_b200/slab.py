"""Latitude-row slab decomposition and the north-south halo exchange.

The tracer path shards by contiguous ranges of latitude rows j: every coupling in j is a
compact stencil (2 rows each side: FCT needs t_lo and R+-Y at j+-1, which reach t at j+-2,
09/mom/tracer_adv_flx.F:554-573,722-739), i is cyclic and stays whole inside a slab, and
columns (invtri, MOBI, convct2) are independent.  Each step only the newly computed level
t(tau+1) crosses the slab boundary: 2 rows x imt x km x nt doubles per side.  One process
per GPU; `torch.distributed` send/recv pairs (NCCL over NVLink on the GPU box, gloo in the
CPU tests) carry the rows -- there is no other data-path collective.  Inventories are
combined with one all-reduce in rank order.
"""
from __future__ import annotations

import numpy as np


def partition_rows(jmt: int, nranks: int):
    """Contiguous (jlo, jhi) per rank covering rows 2..jmt-1, as even as possible."""
    nrows = jmt - 2
    if nranks < 1 or nrows < 2 * nranks:
        raise ValueError(f"need at least 2 rows per slab: jmt={jmt}, nranks={nranks}")
    base, rem = divmod(nrows, nranks)
    out, j = [], 2
    for r in range(nranks):
        n = base + (1 if r < rem else 0)
        out.append((j, j + n - 1))
        j += n
    return out


def partition_rows_balanced(kmt, nranks: int, km: int, land_cost: float = 0.5):
    """Contiguous (jlo, jhi) per rank covering rows 2..jmt-1 with equal estimated WORK instead of equal row counts.

    The kernels skip land (continents, levels below the bottom): a row's cost is its wet cells plus `land_cost` of a
    wet cell for every cell of the row (staging, masks, the parts that are not skipped; fitted to the measured step
    times of slabs with 5 % to 75 % ocean).  With the polar land caps of the synthetic bathymetry equal row counts leave
    the equatorial slabs with twice the mean work; equal work keeps the step time of N slabs at the one-slab time.
    kmt: (jmt, imt) level counts.  Every slab gets at least 2 rows (the halo width)."""
    kmt = np.asarray(kmt)
    jmt, imt = kmt.shape
    nrows = jmt - 2
    if nranks < 1 or nrows < 2 * nranks:
        raise ValueError(f"need at least 2 rows per slab: jmt={jmt}, nranks={nranks}")
    w = kmt[1:-1, 1:-1].sum(axis=1).astype(np.float64) + land_cost * (imt - 2) * km      # rows 2..jmt-1
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, nranks):
        target = cum[-1] * r / nranks
        c = int(np.searchsorted(cum, target))
        c = max(c, cuts[-1] + 2)                       # at least 2 rows in the slab just closed
        c = min(c, nrows - 2 * (nranks - r))           # and room for 2 rows in every slab still to come
        cuts.append(c)
    cuts.append(nrows)
    return [(2 + cuts[r], 2 + cuts[r + 1] - 1) for r in range(nranks)]


def halo_plan(jmt: int, nranks: int, rank: int, parts=None):
    """Local row slices (0-based, into the slab's jl rows) of the four halo pieces.

    Returns dict with keys send_up / recv_up (neighbour rank+1) and send_dn / recv_dn
    (neighbour rank-1); each value is a python slice over the local row axis or None."""
    parts = parts if parts is not None else partition_rows(jmt, nranks)
    jlo, jhi = parts[rank]
    jbase = max(1, jlo - 2)
    plan = dict(send_up=None, recv_up=None, send_dn=None, recv_dn=None, jlo=jlo, jhi=jhi, jbase=jbase)
    if rank + 1 < nranks:
        plan["send_up"] = slice(jhi - 1 - jbase, jhi + 1 - jbase)      # owned rows jhi-1, jhi
        plan["recv_up"] = slice(jhi + 1 - jbase, jhi + 3 - jbase)      # halo rows jhi+1, jhi+2
    if rank > 0:
        plan["send_dn"] = slice(jlo - jbase, jlo + 2 - jbase)          # owned rows jlo, jlo+1
        plan["recv_dn"] = slice(jlo - 2 - jbase, jlo - jbase)          # halo rows jlo-2, jlo-1
    return plan


class HaloExchanger:
    """Exchanges the 2-row halos of a (nt, jl, km, imt) torch tensor between neighbouring ranks."""

    def __init__(self, jmt, rank, world, dist=None, parts=None):
        self.rank, self.world = rank, world
        self.plan = halo_plan(jmt, world, rank, parts)
        self.dist = dist
        self._bufs = {}

    def exchange(self, t):
        """t: torch tensor (nt, jl, km, imt), CUDA (nccl) or CPU (gloo); halos updated in place."""
        if self.world == 1:
            return
        import torch

        dist = self.dist or torch.distributed
        p = self.plan
        ops, recvs = [], []
        for key_s, key_r, peer in (("send_up", "recv_up", self.rank + 1), ("send_dn", "recv_dn", self.rank - 1)):
            if p[key_s] is None:
                continue
            sbuf = t[:, p[key_s]].contiguous()
            rbuf = self._bufs.get(key_r)
            if rbuf is None or rbuf.shape != sbuf.shape or rbuf.device != sbuf.device:
                rbuf = torch.empty_like(sbuf)
                self._bufs[key_r] = rbuf
            ops.append(dist.P2POp(dist.isend, sbuf, peer))
            ops.append(dist.P2POp(dist.irecv, rbuf, peer))
            recvs.append((key_r, rbuf))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for key_r, rbuf in recvs:
            t[:, p[key_r]].copy_(rbuf)

    def exchange_async(self, t, main_stream, side_stream):
        """The same exchange on `side_stream`, behind everything `main_stream` holds so far; returns a torch.cuda.Event
        that fires when the halo rows are in place (None with one rank).  The caller keeps launching on `main_stream`
        and makes the first reader of the new halo rows wait for the event (uvic_b200_wait_before_advection)."""
        if self.world == 1:
            return None
        import torch

        side_stream.wait_stream(main_stream)
        with torch.cuda.stream(side_stream):
            self.exchange(t)
            ev = torch.cuda.Event()
            ev.record(side_stream)
        return ev

    def bytes_per_step(self, nt, km, imt):
        n = sum(1 for k in ("send_up", "send_dn") if self.plan[k] is not None)
        return n * 2 * imt * km * nt * 8


def device_tensor(ptr: int, shape, device_index: int):
    """Zero-copy torch view of library-owned device memory (float64, C order)."""
    import torch

    class _Arr:
        pass

    a = _Arr()
    a.__cuda_array_interface__ = {
        "shape": tuple(int(s) for s in shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2, "strides": None,
    }
    return torch.as_tensor(a, device=f"cuda:{device_index}")


def combine_inventories(local_inv: np.ndarray, dist=None, device=None):
    """Global tracer inventories: gather the per-slab partial sums and add them in rank
    order (fixed order -> reproducible to the bit for a given decomposition)."""
    import torch

    dist = dist or torch.distributed
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_inv.copy()
    x = torch.from_numpy(np.ascontiguousarray(local_inv))
    if device is not None:
        x = x.to(device)
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, x)
    tot = parts[0].clone()
    for q in parts[1:]:
        tot += q
    return tot.cpu().numpy()
