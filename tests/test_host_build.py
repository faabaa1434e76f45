"""CPU: the PRODUCT's step arithmetic and noise (soilmachine_b200/csrc/sm_core.cuh, sm_noise.cuh),
compiled for the host by tests/hostsim, against the golden vectors - this pins the type-exact
transcription the CUDA kernels execute without needing a GPU."""
import numpy as np
import pytest
import _golden
import _hostsim


class Backend:
    """adapts tests/hostsim to the driver interface of _golden.replay_frame"""

    def __init__(self, g):
        self.hs = _hostsim.HostSim()
        self.hs.init(int(g["dimx"]), int(g["dimy"]), int(g["scale"]), g["soils"])

    def columns(self):
        return self.hs.columns()

    def heights(self):
        return self.hs.heights()

    def frequency(self):
        return self.hs.frequency()

    def frequency_update(self):
        f = self.hs.frequency()
        lrate, K = np.float32(0.01), np.float32(50.0)
        t = f["water_track"]
        one = np.float32(1.0)
        new = (one - lrate) * f["water_frequency"] + lrate * K * t / (one + K * t)
        self.hs.set_frequency(new.astype(np.float32), np.zeros_like(t), None)

    def water_run(self, xy):
        st = _hostsim.Stats()
        self.hs.water_begin(xy)
        while self.hs.water_sweep(st):
            pass
        return st

    def wind_run(self, xy):
        st = _hostsim.Stats()
        self.hs.wind_begin(xy)
        while self.hs.wind_sweep(st):
            pass
        return st

    def water_state(self):
        return self.hs.water_state()

    def wind_state(self):
        return self.hs.wind_state()


def stats5(st):
    return (st.steps, st.sweeps, st.exit_oob, st.exit_evap, st.exit_stall)


@pytest.mark.parametrize("case", _golden.FRAME_CASES)
def test_product_arithmetic_replays_golden_frame(case):
    g = _golden.load(case)
    b = Backend(g)
    b.hs.set_columns(_golden.cols(g, "init"))
    _golden.replay_frame(g, b, stats5)


@pytest.mark.parametrize("case", _golden.FRAME_CASES)
def test_product_noise_reproduces_initial_terrain(case):
    """sm_noise.cuh (OpenSimplex2/FBm restatement) == Layermap::initialize of the reference."""
    g = _golden.load(case)
    b = Backend(g)
    b.hs.initialize(int(g["seed"]), g["layers"])
    _golden.same_cols(b.hs.columns(), _golden.cols(g, "init"), "initial terrain")


def test_product_column_ops_truth_table():
    g = _golden.load("column_ops")
    hs = _hostsim.HostSim()
    hs.init(8, 8, int(g["scale"]), g["soils"])
    for (kind, x, y, v, t), want in zip(g["ops"], g["remove_results"]):
        if kind == 0:
            hs.lib.hs_add(int(x), int(y), float(v), int(t))
        else:
            got = hs.lib.hs_remove(int(x), int(y), float(v))
            assert np.float64(got).tobytes() == np.float64(want).tobytes()
    _golden.same_cols(hs.columns(), _golden.cols(g, "final"), "columns after add/remove")
    normals = np.array([hs.normal(x, y) for x in range(8) for y in range(8)], np.float32)
    _golden.same(normals, g["normals"], "normals")
    bil = np.array([hs.lib.hs_height_f(float(p[0]), float(p[1])) for p in g["bilinear_pts"]])
    _golden.same(bil, g["bilinear"], "bilinear heights")
    for x, y, loop in g["cascades"]:
        hs.lib.hs_cascade(float(x), float(y), int(loop))
    _golden.same_cols(hs.columns(), _golden.cols(g, "after_cascade"), "columns after cascades")
