"""ctypes binding of oracle/liblbmoracle.so (oracle/lbm_oracle.c: CPU restatement of the reference's LBM wind
shaders).  TEST INFRASTRUCTURE ONLY - parity with upstream's GLSL is unpinned, see the C file's header."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblbmoracle.so")


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class Lbm:
    def __init__(self, nx, ny, nz):
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "lbm_oracle.c")):
            subprocess.check_call(["make", "-C", HERE, "lbm"], stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(LIB)
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.n = self.nx * self.ny * self.nz
        self.lib.lbmo_create(self.nx, self.ny, self.nz)
        self.lib.lbmo_init()                       # lbmwind.h:101-107: init.cs runs with an all-zero boundary

    def set_boundary(self, b):
        b = np.ascontiguousarray(b, np.float32)
        assert b.size == self.n
        self.lib.lbmo_set_boundary(_p(b))

    def init(self):
        self.lib.lbmo_init()

    def step(self, n=1):
        self.lib.lbmo_step(int(n))

    def get(self):
        f = np.zeros((self.n, 19), np.float32); rho = np.zeros(self.n, np.float32); v = np.zeros((self.n, 4), np.float32)
        self.lib.lbmo_get(_p(f), _p(rho), _p(v))
        return {"f": f, "rho": rho, "v": v}

    def advect(self, pos4):
        pos = np.ascontiguousarray(pos4, np.float32).copy()
        self.lib.lbmo_advect(len(pos), _p(pos))
        return pos
