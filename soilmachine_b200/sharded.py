"""Sharded maps: one sm_context per rank, x-strips, peer pointers.

VirtualShards  - all ranks are contexts of THIS process on one GPU (share = nranks).  Same kernels, same
                 cross-rank barrier and hand-over protocol as the multi-GPU path, exercised on a single
                 device (what the 1-GPU test tier runs).
DistShard      - one rank per process / GPU under torch.distributed: peer blobs are all-gathered and the
                 other GPUs' arrays are opened with CUDA IPC, so halo records, bins and `done` words are
                 read and written over NVLink by the sweep kernel itself.
"""
import ctypes as C
import numpy as np
from . import capi


def split_columns(cols, dimx, dimy, x0, x1):
    """global bottom->top CSR (cell order x*dimy+y) -> the CSR of the strip [x0, x1)."""
    off = np.asarray(cols["offsets"])
    lo, hi = off[x0 * dimy], off[x1 * dimy]
    out = {"offsets": (off[x0 * dimy:x1 * dimy + 1] - lo).astype(np.int64)}
    for k in ("type", "size", "saturation"):
        out[k] = None if cols.get(k) is None else np.asarray(cols[k])[lo:hi]
    return out


def merge_columns(parts):
    off = [np.zeros(1, np.int64)]
    base = 0
    for p in parts:
        off.append(p["offsets"][1:] + base)
        base += p["offsets"][-1]
    out = {"offsets": np.concatenate(off)}
    for k in ("type", "size", "floor", "saturation"):
        out[k] = np.concatenate([p[k] for p in parts])
    return out


def merge_frequency(parts, ranges, dimx, dimy):
    """frequency arrays are indexed y*dimx + x; rank q's copy is authoritative for its own columns."""
    out = {}
    for k in parts[0]:
        m = np.zeros((dimy, dimx), np.float32)
        for p, (x0, x1) in zip(parts, ranges):
            m[:, x0:x1] = p[k].reshape(dimy, dimx)[:, x0:x1]
        out[k] = m.reshape(-1)
    return out


def sum_stats(stats):
    tot = capi.Stats()
    for f, _ in capi.Stats._fields_:
        vals = [getattr(s, f) for s in stats]
        setattr(tot, f, max(vals) if f in ("sweeps", "device_ms", "alive") else sum(vals))
    return tot


class VirtualShards:
    def __init__(self, nranks, dimx, dimy, scale=80, device=0, max_particles=0, pool_capacity=0, budget=False):
        self.nranks, self.dimx, self.dimy = nranks, dimx, dimy
        self.ctx = [capi.Context(dimx, dimy, scale, device=device, max_particles=max_particles,
                                 pool_capacity=pool_capacity, nranks=nranks, rank=r, share=nranks, budget=budget)
                    for r in range(nranks)]
        blobs = [c.peer_export() for c in self.ctx]
        for c in self.ctx:
            c.peer_attach(blobs, use_ipc=False)
        self.ranges = [(c.x0, c.x1) for c in self.ctx]

    def close(self):
        for c in self.ctx:
            c.close()

    def set_soils(self, table):
        for c in self.ctx:
            c.set_soils(table)

    def initialize(self, seed, layers):
        for c in self.ctx:
            c.initialize(seed, layers)

    def upload_columns(self, cols):
        for c in self.ctx:
            p = split_columns(cols, self.dimx, self.dimy, c.x0, c.x1)
            c.upload_columns(p["offsets"], p["type"], p["size"], p["saturation"])

    def download_columns(self):
        return merge_columns([c.download_columns() for c in self.ctx])

    def heights(self):
        return np.concatenate([c.heights() for c in self.ctx], axis=0)

    def frequency(self):
        return merge_frequency([c.frequency() for c in self.ctx], self.ranges, self.dimx, self.dimy)

    def frequency_update(self):
        for c in self.ctx:
            c.frequency_update()

    def _run(self, kind, xy, max_sweeps):
        xy = np.ascontiguousarray(xy, np.float32)
        spawn = [c.device_spawn(xy) if len(xy) else None for c in self.ctx]
        for c, d in zip(self.ctx, spawn):          # launch every rank before waiting for any
            (c.water_run_device if kind == "water" else c.wind_run_device)(d, len(xy), max_sweeps)
        stats = [c.last_stats() for c in self.ctx]
        for c, d in zip(self.ctx, spawn):
            if d is not None:
                c.device_free(d)
        return sum_stats(stats)

    def water_run(self, xy, max_sweeps=0):
        return self._run("water", xy, max_sweeps)

    def wind_run(self, xy, max_sweeps=0):
        return self._run("wind", xy, max_sweeps)


class DistShard:
    """This process's rank of a map sharded over torch.distributed ranks (one GPU each)."""

    def __init__(self, dimx, dimy, scale, device, max_particles=0, pool_capacity=0, share=1):
        """share > 1: that many ranks (processes) run their kernels on the SAME device - used by the one-GPU
        test of the CUDA-IPC path; every rank's grid must then be resident at the same time."""
        import torch.distributed as dist
        self.dist = dist
        self.nranks, self.rank = dist.get_world_size(), dist.get_rank()
        self.dimx, self.dimy = dimx, dimy
        self.ctx = capi.Context(dimx, dimy, scale, device=device, max_particles=max_particles,
                                pool_capacity=pool_capacity, nranks=self.nranks, rank=self.rank, share=share)
        mine = bytes(self.ctx.peer_export())
        blobs = [None] * self.nranks
        dist.all_gather_object(blobs, mine)
        self.ctx.peer_attach([capi.PeerBlob.from_buffer_copy(b) for b in blobs], use_ipc=True)
        dist.barrier()

    def run(self, kind, d_xy, n, max_sweeps=0):
        """launch this rank's sweep kernel (all ranks must call it), wait, return local stats"""
        (self.ctx.water_run_device if kind == "water" else self.ctx.wind_run_device)(d_xy, n, max_sweeps)
        return self.ctx.last_stats()

    def close(self):
        self.dist.barrier()
        self.ctx.close()
