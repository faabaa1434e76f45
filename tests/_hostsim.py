"""ctypes binding of tests/hostsim/libhostsim.so (the product's step arithmetic compiled for the
host; test tool only)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
LIB = os.path.join(HERE, "hostsim", "libhostsim.so")
CORE = os.path.join(HERE, "..", "soilmachine_b200", "csrc", "sm_core.cuh")
NOISE = os.path.join(HERE, "..", "soilmachine_b200", "csrc", "sm_noise.cuh")
HYDRO = os.path.join(HERE, "..", "soilmachine_b200", "csrc", "sm_hydro.cuh")
COOP = os.path.join(HERE, "..", "soilmachine_b200", "csrc", "sm_coop.cuh")
HCOOP = os.path.join(HERE, "..", "soilmachine_b200", "csrc", "sm_hydro_coop.cuh")

SOILDEV = np.dtype([("friction", "<f4"), ("solubility", "<f4"), ("equrate", "<f4"), ("erosionrate", "<f4"),
                    ("maxdiff", "<f4"), ("settling", "<f4"), ("suspension", "<f4"), ("porosity", "<f4"),
                    ("transports", "<u4"), ("erodes", "<u4"), ("cascades", "<u4"), ("abrades", "<u4")])


class Stats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("sweeps", C.c_int64), ("exit_oob", C.c_int64),
                ("exit_evap", C.c_int64), ("exit_stall", C.c_int64), ("seconds", C.c_double)]


class HydroCount(C.Structure):
    _fields_ = [("floods", C.c_uint64), ("nested", C.c_uint64), ("nested_steps", C.c_uint64),
                ("transfers", C.c_uint64), ("cells", C.c_uint64), ("overflow", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build():
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(NOISE), os.path.getmtime(HYDRO), os.path.getmtime(COOP), os.path.getmtime(HCOOP)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", LIB])
    return LIB


def soildev_from(soils):
    out = np.zeros(len(soils), SOILDEV)
    for k in SOILDEV.names:
        out[k] = soils[k]
    return out


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class HostSim:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.hs_nsections.restype = C.c_int64
        L.hs_height_f.restype = C.c_double
        L.hs_height_f.argtypes = [C.c_float, C.c_float]
        L.hs_remove.restype = C.c_double
        L.hs_remove.argtypes = [C.c_int, C.c_int, C.c_double]
        L.hs_add.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.hs_cascade.argtypes = [C.c_float, C.c_float, C.c_int]

    def init(self, dimx, dimy, scale, soils):
        self.dimx, self.dimy = dimx, dimy
        sd = soildev_from(soils)
        self.lib.hs_init(dimx, dimy, scale, len(sd), sd.ctypes.data_as(C.c_void_p))

    def initialize(self, seed, layers):
        lay = np.ascontiguousarray(layers)
        self.lib.hs_initialize(int(seed), len(lay), lay.ctypes.data_as(C.c_void_p))

    def set_columns(self, cols):
        off = np.ascontiguousarray(cols["offsets"], np.int64)
        typ = np.ascontiguousarray(cols["type"], np.int32)
        size = np.ascontiguousarray(cols["size"], np.float64)
        sat = np.ascontiguousarray(cols["saturation"], np.float64)
        self.lib.hs_set_columns(_p(off, C.c_int64), _p(typ, C.c_int32), _p(size, C.c_double), _p(sat, C.c_double))

    def columns(self):
        n = self.lib.hs_nsections()
        cells = self.dimx * self.dimy
        off = np.zeros(cells + 1, np.int64); typ = np.zeros(n, np.int32)
        size = np.zeros(n); floor = np.zeros(n); sat = np.zeros(n)
        self.lib.hs_get_columns(_p(off, C.c_int64), _p(typ, C.c_int32), _p(size, C.c_double),
                                _p(floor, C.c_double), _p(sat, C.c_double))
        return {"offsets": off, "type": typ, "size": size, "floor": floor, "saturation": sat}

    def heights(self):
        out = np.zeros(self.dimx * self.dimy)
        self.lib.hs_heights(_p(out, C.c_double))
        return out.reshape(self.dimx, self.dimy)

    def frequency(self):
        a = [np.zeros(self.dimx * self.dimy, np.float32) for _ in range(3)]
        self.lib.hs_get_frequency(*[_p(x, C.c_float) for x in a])
        return {"water_frequency": a[0], "water_track": a[1], "wind_frequency": a[2]}

    def set_frequency(self, water_frequency=None, water_track=None, wind_frequency=None):
        arrs = [None if x is None else np.ascontiguousarray(x, np.float32)
                for x in (water_frequency, water_track, wind_frequency)]
        self.lib.hs_set_frequency(*[_p(x, C.c_float) for x in arrs])

    def normal(self, x, y):
        o = (C.c_float * 3)()
        self.lib.hs_normal(int(x), int(y), o)
        return np.array(list(o), np.float32)

    def set_wind_field(self, v4=None, dims=None):
        if v4 is None:
            self.lib.hs_set_wind_field(None, 0, 0, 0)
            return
        v = np.ascontiguousarray(v4, np.float32)
        self.lib.hs_set_wind_field(_p(v, C.c_float), int(dims[0]), int(dims[1]), int(dims[2]))

    def budget(self):
        """coop mode: per-particle mass-budget accumulators [n, 6] of the current batch"""
        per = np.zeros((self._n, 6))
        self.lib.hs_budget(_p(per, C.c_double))
        return per

    def water_begin(self, xy):
        xy = np.ascontiguousarray(xy, np.float32); self._n = len(xy)
        self.lib.hs_water_begin(len(xy), _p(xy, C.c_float))

    def water_sweep(self, st):
        return self.lib.hs_water_sweep(C.byref(st))

    def water_state(self):
        n = self._n
        pos = np.zeros((n, 2), np.float32); speed = np.zeros((n, 2), np.float32)
        vol = np.zeros(n); sed = np.zeros(n); cont = np.zeros(n, np.int32); alive = np.zeros(n, np.int32)
        self.lib.hs_water_state(_p(pos, C.c_float), _p(speed, C.c_float), _p(vol, C.c_double),
                                _p(sed, C.c_double), _p(cont, C.c_int32), _p(alive, C.c_int32))
        return {"pos": pos, "speed": speed, "volume": vol, "sediment": sed, "contains": cont, "alive": alive}

    def water_run(self, xy):
        self.water_begin(xy)
        st = Stats()
        while self.water_sweep(st) > 0:
            pass
        return st

    def water_flood(self):
        hc = HydroCount()
        self.lib.hs_water_flood(C.byref(hc))
        return hc

    def seep(self, mode=1):
        hc = HydroCount()
        self.lib.hs_seep(int(mode), C.byref(hc))
        return hc

    def frequency_update(self):
        f = self.frequency()
        lrate, K = np.float32(0.01), np.float32(50.0)
        t = f["water_track"]
        wf = (np.float32(1.0) - lrate) * f["water_frequency"] + lrate * K * t / (np.float32(1.0) + K * t)
        self.set_frequency(water_frequency=wf.astype(np.float32), water_track=np.zeros_like(t))

    def wind_begin(self, xy):
        xy = np.ascontiguousarray(xy, np.float32); self._n = len(xy)
        self.lib.hs_wind_begin(len(xy), _p(xy, C.c_float))

    def wind_sweep(self, st):
        return self.lib.hs_wind_sweep(C.byref(st))

    def wind_state(self):
        n = self._n
        pos = np.zeros((n, 2), np.float32); speed = np.zeros((n, 3), np.float32)
        h = np.zeros(n); sed = np.zeros(n); cont = np.zeros(n, np.int32); alive = np.zeros(n, np.int32)
        self.lib.hs_wind_state(_p(pos, C.c_float), _p(speed, C.c_float), _p(h, C.c_double),
                               _p(sed, C.c_double), _p(cont, C.c_int32), _p(alive, C.c_int32))
        return {"pos": pos, "speed": speed, "height": h, "sediment": sed, "contains": cont, "alive": alive}
