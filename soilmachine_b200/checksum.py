"""The column checksum of sm_checksum (soilmachine_b200/csrc/sm_engine.cu: k_checksum) in numpy.

Every section contributes mix(cell, depth from the top, size, floor, saturation, type); contributions are
summed modulo 2^64.  `cell` is the GLOBAL cell index x*dimy + y, so the checksums of the x-strips of a sharded
map add up to the checksum of the whole map.  Used by the parity tests (reference columns vs device) and by
bench.py (equal checksums on 1, 2, 4, 8 GPUs).
"""
import numpy as np

_M = (1 << 64) - 1


def _mix64(z):
    z = z ^ (z >> np.uint64(30)); z = z * np.uint64(0xbf58476d1ce4e5b9)
    z = z ^ (z >> np.uint64(27)); z = z * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))


def columns_checksum(cols, first_cell=0):
    """cols = bottom->top CSR {"offsets", "type", "size", "floor", "saturation"} as downloaded from a context
    or from the reference; first_cell = global index of the CSR's first cell (x0*dimy for a strip)."""
    off = np.asarray(cols["offsets"], np.int64)
    n = int(off[-1])
    if n == 0:
        return 0
    counts = np.diff(off)
    cell = np.repeat(np.arange(len(counts), dtype=np.uint64) + np.uint64(first_cell), counts)
    top = np.repeat(off[1:] - 1, counts)                       # index of each column's top section
    depth = (top - np.arange(n, dtype=np.int64)).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = _mix64(cell * np.uint64(0x9e3779b97f4a7c15) + depth)
        for k in ("size", "floor", "saturation"):
            h = _mix64(h ^ np.ascontiguousarray(cols[k], np.float64).view(np.uint64))
        h = _mix64(h ^ np.asarray(cols["type"]).astype(np.int64).astype(np.uint64))
        return int(np.add.reduce(h, dtype=np.uint64)) & _M
