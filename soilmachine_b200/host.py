"""Host-side mirror of the GL-free part of the reference's main() (SoilMachine.cpp:34-48,82-83,
283-320) on top of the C ABI: load a preset, build the terrain on the GPU, run frames.

A frame = one water batch run to completion, optionally the pooling hydrology (the batch's floods and
the full-grid seep pass, SoilMachine.cpp:296,300-301), one wind batch run to completion, the frequency
update.  Spawn positions are drawn with
the C library's rand() in the order the particle constructors draw them (water.h:13, wind.h:15:
GCC evaluates the two arguments right to left, so y takes the first draw).
"""
import ctypes as C
import ctypes.util
import os
import numpy as np
from . import capi, presets

_libc = None


def _c():
    global _libc
    if _libc is None:
        _libc = C.CDLL(ctypes.util.find_library("c") or "libc.so.6")
        _libc.rand.restype = C.c_int
        _libc.srand.argtypes = [C.c_uint]
    return _libc


def srand(seed):
    _c().srand(int(seed) & 0xFFFFFFFF)


def spawn_list(n, dimx, dimy):
    """n (x, y) spawn positions = n constructor calls' worth of rand() draws."""
    lc = _c()
    out = np.empty((n, 2), np.float32)
    for i in range(n):
        y = lc.rand() % dimy
        x = lc.rand() % dimx
        out[i, 0] = x
        out[i, 1] = y
    return out


class Simulation:
    """soil preset + GPU context + frame loop."""

    def __init__(self, soil, seed=42, dimx=0, dimy=0, device=0, max_particles=0, pool_capacity=0):
        # a preset name (JSON tables dumped from the reference loader) or a path to a `.soil` file
        self.preset = capi.parse_soil_file(soil) if str(soil).endswith(".soil") and os.path.exists(str(soil)) \
            else presets.load(soil)
        w = self.preset["world"]
        self.dimx = int(dimx or w["sizex"])
        self.dimy = int(dimy or w["sizey"])
        self.scale = int(w["scale"])
        self.seed = int(seed)
        srand(self.seed)                                   # SoilMachine.cpp:41
        self.ctx = capi.Context(self.dimx, self.dimy, self.scale, device=device,
                                pool_capacity=pool_capacity, max_particles=max_particles)
        self.ctx.set_soils(self.preset["soils"])
        self.ctx.set_soil_colors(self.preset["colors"])
        self.ctx.initialize(self.seed, self.preset["layers"])   # Layermap(SEED, dim), SoilMachine.cpp:83

    def frame(self, nwater, nwind, water_xy=None, wind_xy=None, hydrology=False, water_chunk=0):
        """SoilMachine.cpp:287-320.  hydrology=False: the hot path only (flood disabled, no seep pass);
        True: the water batch is followed by its floods and by the seep pass (self.last_hydrology holds the
        counter sets).  water_chunk > 0 splits the frame's water particles into lockstep batches of that
        size, each followed by its floods: upstream floods every particle right after its own loop, so later
        particles of a frame meet the ponds earlier ones left; smaller chunks follow that interleaving more
        closely (chunk 1 = upstream's order) at the price of fewer particles in flight.
        Returns (water_stats of the last batch, wind_stats)."""
        ws = ds = None
        self.last_hydrology = None
        if nwater:
            xy = water_xy if water_xy is not None else spawn_list(nwater, self.dimx, self.dimy)
            chunk = int(water_chunk) if (hydrology and water_chunk and water_chunk > 0) else len(xy)
            floods = []
            for i in range(0, len(xy), chunk):
                ws = self.ctx.water_run(xy[i:i + chunk])
                if hydrology:
                    floods.append(self.ctx.water_flood())
            if hydrology:
                self.last_hydrology = (floods, self.ctx.seep())
        if nwind:
            xy = wind_xy if wind_xy is not None else spawn_list(nwind, self.dimx, self.dimy)
            ds = self.ctx.wind_run(xy)
        if nwater:
            self.ctx.frequency_update()
        return ws, ds

    def close(self):
        self.ctx.close()
