"""ctypes binding of oracle/libsmoracle.so -- the plain CPU restatement (oracle/sm_oracle.cpp).
TEST INFRASTRUCTURE: same driver interface as oracle/refapi.py so tests can swap one for the other."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsmoracle.so")

SOIL_DTYPE = np.dtype([
    ("transports", "<i4"), ("erodes", "<i4"), ("cascades", "<i4"), ("abrades", "<i4"),
    ("density", "<f4"), ("porosity", "<f4"), ("solubility", "<f4"), ("equrate", "<f4"),
    ("friction", "<f4"), ("erosionrate", "<f4"), ("maxdiff", "<f4"), ("settling", "<f4"),
    ("suspension", "<f4"), ("abrasion", "<f4"),
])


class Stats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("sweeps", C.c_int64), ("exit_oob", C.c_int64),
                ("exit_evap", C.c_int64), ("exit_stall", C.c_int64), ("seconds", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Hydro(C.Structure):
    _fields_ = [("floods", C.c_int64), ("nested", C.c_int64), ("nested_steps", C.c_int64),
                ("transfers", C.c_int64), ("cells", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build():
    src = [os.path.join(HERE, "sm_oracle.cpp"), os.path.join(HERE, "sm_oracle.h")]
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Port:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.smo_nsections.restype = C.c_int64
        L.smo_height_i.restype = C.c_double
        L.smo_height_i.argtypes = [C.c_int, C.c_int]
        L.smo_height_f.restype = C.c_double
        L.smo_height_f.argtypes = [C.c_float, C.c_float]
        L.smo_surface.argtypes = [C.c_int, C.c_int]
        L.smo_normal.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.smo_add.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.smo_remove.restype = C.c_double
        L.smo_remove.argtypes = [C.c_int, C.c_int, C.c_double]
        L.smo_cascade.argtypes = [C.c_float, C.c_float, C.c_int]

    def init(self, dimx, dimy, scale, soils):
        self.dimx, self.dimy, self.scale = int(dimx), int(dimy), int(scale)
        s = np.zeros(len(soils), SOIL_DTYPE)
        for k in SOIL_DTYPE.names:
            s[k] = soils[k]
        self.lib.smo_init(self.dimx, self.dimy, self.scale, len(s), s.ctypes.data_as(C.c_void_p))
        return self

    @property
    def cells(self):
        return self.dimx * self.dimy

    def set_columns(self, cols):
        off = np.ascontiguousarray(cols["offsets"], np.int64)
        typ = np.ascontiguousarray(cols["type"], np.int32)
        size = np.ascontiguousarray(cols["size"], np.float64)
        sat = np.ascontiguousarray(cols["saturation"], np.float64) if cols.get("saturation") is not None else None
        self.lib.smo_set_columns(_p(off, C.c_int64), _p(typ, C.c_int32), _p(size, C.c_double), _p(sat, C.c_double))

    def columns(self):
        n = self.lib.smo_nsections()
        off = np.zeros(self.cells + 1, np.int64); typ = np.zeros(n, np.int32)
        size = np.zeros(n); floor = np.zeros(n); sat = np.zeros(n)
        self.lib.smo_get_columns(_p(off, C.c_int64), _p(typ, C.c_int32), _p(size, C.c_double),
                                 _p(floor, C.c_double), _p(sat, C.c_double))
        return {"offsets": off, "type": typ, "size": size, "floor": floor, "saturation": sat}

    def heights(self):
        out = np.zeros(self.cells)
        self.lib.smo_heights(_p(out, C.c_double))
        return out.reshape(self.dimx, self.dimy)

    def frequency(self):
        a = [np.zeros(self.cells, np.float32) for _ in range(3)]
        self.lib.smo_get_frequency(*[_p(x, C.c_float) for x in a])
        return {"water_frequency": a[0], "water_track": a[1], "wind_frequency": a[2]}

    def set_frequency(self, water_frequency=None, water_track=None, wind_frequency=None):
        arrs = [None if x is None else np.ascontiguousarray(x, np.float32)
                for x in (water_frequency, water_track, wind_frequency)]
        self.lib.smo_set_frequency(*[_p(x, C.c_float) for x in arrs])

    def frequency_update(self):
        self.lib.smo_frequency_update()

    def height(self, x, y):
        if isinstance(x, (int, np.integer)) and isinstance(y, (int, np.integer)):
            return self.lib.smo_height_i(int(x), int(y))
        return self.lib.smo_height_f(float(x), float(y))

    def surface(self, x, y):
        return self.lib.smo_surface(int(x), int(y))

    def normal(self, x, y):
        o = (C.c_float * 3)()
        self.lib.smo_normal(int(x), int(y), o)
        return np.array(list(o), np.float32)

    def add(self, x, y, size, typ):
        self.lib.smo_add(int(x), int(y), float(size), int(typ))

    def remove(self, x, y, h):
        return self.lib.smo_remove(int(x), int(y), float(h))

    def cascade(self, x, y, transferloop=0):
        self.lib.smo_cascade(float(x), float(y), int(transferloop))

    def _run(self, fn, xy, max_sweeps=None):
        xy = np.ascontiguousarray(xy, np.float32)
        st = Stats()
        if max_sweeps is None:
            fn(len(xy), _p(xy, C.c_float), C.byref(st))
        else:
            fn(len(xy), _p(xy, C.c_float), int(max_sweeps), C.byref(st))
        return st

    def water_run(self, xy, max_sweeps=0):
        self._nw = len(xy)
        return self._run(self.lib.smo_water_run, xy, max_sweeps)

    def wind_run(self, xy, max_sweeps=0):
        self._nd = len(xy)
        return self._run(self.lib.smo_wind_run, xy, max_sweeps)

    def set_wind_field(self, v4=None, dims=None):
        """attach a lattice velocity field [nx*ny*nz, 4] for the wind particles' prevailing wind (None: detach)"""
        if v4 is None:
            self.lib.smo_set_wind_field(None, 0, 0, 0)
            return
        v = np.ascontiguousarray(v4, np.float32)
        self.lib.smo_set_wind_field(_p(v, C.c_float), int(dims[0]), int(dims[1]), int(dims[2]))

    def budget(self):
        """mass budget of the last lockstep batch: (per-particle accumulators [n, 6], their sums in particle order
        [6]) - eroded, deposited, cascade_net, discarded, clamped, wind_negative (sm_coop.cuh)"""
        self.lib.smo_budget.restype = C.c_int64
        n = int(self.lib.smo_budget(None, None))
        per = np.zeros((n, 6)); sums = np.zeros(6)
        self.lib.smo_budget(_p(per, C.c_double), _p(sums, C.c_double))
        return per, sums

    def water_seq(self, xy):
        return self._run(self.lib.smo_water_seq, xy)

    def water_seq_full(self, xy, flood=True, seep=True):
        xy = np.ascontiguousarray(xy, np.float32)
        st, hy = Stats(), Hydro()
        self.lib.smo_water_seq_full(len(xy), _p(xy, C.c_float), int(flood) | (int(seep) << 1), C.byref(st), C.byref(hy))
        return st, hy

    def water_flood(self):
        hy = Hydro()
        self.lib.smo_water_flood(C.byref(hy))
        return hy

    def seep(self):
        hy = Hydro()
        self.lib.smo_seep(C.byref(hy))
        return hy

    def wind_seq(self, xy):
        return self._run(self.lib.smo_wind_seq, xy)

    def water_state(self):
        n = self._nw
        pos = np.zeros((n, 2), np.float32); speed = np.zeros((n, 2), np.float32)
        vol = np.zeros(n); sed = np.zeros(n); cont = np.zeros(n, np.int32); alive = np.zeros(n, np.int32)
        self.lib.smo_water_state(_p(pos, C.c_float), _p(speed, C.c_float), _p(vol, C.c_double),
                                 _p(sed, C.c_double), _p(cont, C.c_int32), _p(alive, C.c_int32))
        return {"pos": pos, "speed": speed, "volume": vol, "sediment": sed, "contains": cont, "alive": alive}

    def wind_state(self):
        n = self._nd
        pos = np.zeros((n, 2), np.float32); speed = np.zeros((n, 3), np.float32)
        h = np.zeros(n); sed = np.zeros(n); cont = np.zeros(n, np.int32); alive = np.zeros(n, np.int32)
        self.lib.smo_wind_state(_p(pos, C.c_float), _p(speed, C.c_float), _p(h, C.c_double),
                                _p(sed, C.c_double), _p(cont, C.c_int32), _p(alive, C.c_int32))
        return {"pos": pos, "speed": speed, "height": h, "sediment": sed, "contains": cont, "alive": alive}
