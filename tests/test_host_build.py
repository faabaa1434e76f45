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
def test_warp_cooperative_step_with_late_corner_staging(case):
    """the water step with only the plus-shaped stencil staged before move() and the corners of the first block staged
    with the second block (what the exact-footprint schedule needs, CoopWin::begin / target)"""
    g = _golden.load(case)
    b = Backend(g)
    b.hs.lib.hs_set_mode(2, 0)
    try:
        b.hs.set_columns(_golden.cols(g, "init"))
        _golden.replay_frame(g, b, stats5)
    finally:
        b.hs.lib.hs_set_mode(0, 0)


@pytest.mark.parametrize("case", _golden.FRAME_CASES)
def test_exact_schedule_is_sound_under_the_most_out_of_order_legal_execution(case):
    """Which legal order a GPU run of the exact-footprint schedule takes depends on timing.  tests/hostsim mode 3 runs
    every sweep of the golden frames in the legal order that departs MOST from index order - highest index first, each
    particle advancing a phase (move / interact) as soon as no unfinished lower-index particle's (possible) writes or
    footprint can meet what the phase touches, its staged window living across the two phases while others run in
    between - and must still reproduce the reference byte for byte.  The rule is evaluated on explicit cell sets."""
    g = _golden.load(case)
    b = Backend(g)
    b.hs.lib.hs_adversarial_ahead.restype = __import__("ctypes").c_longlong
    b.hs.lib.hs_set_mode(3, 0)
    try:
        b.hs.lib.hs_adversarial_ahead()
        b.hs.set_columns(_golden.cols(g, "init"))
        _golden.replay_frame(g, b, stats5)
        ahead = b.hs.lib.hs_adversarial_ahead()
    finally:
        b.hs.lib.hs_set_mode(0, 0)
    assert ahead > 1000        # the order really was far from index order


def test_footprint_predicates_equal_the_cell_sets():
    """sm_foot.cuh (the closed forms sweep_exact / k_run_exact evaluate on the device) against explicit cell sets, over
    every relative position that can occur: never "no" when the sets meet (sound), and - except where a predicate is
    documented as a bound - never "yes" when they do not (exact)."""
    import ctypes as C
    from _hostsim import HostSim
    hs = HostSim()
    out = (C.c_longlong * 16)()
    hs.lib.hs_check_foot_predicates(out)
    v = list(out)
    names = ["box_hits_M", "W_hits_M", "F_hits_F", "box_hits_F"]
    for kind in (0, 1):
        for k, nm in enumerate(names):
            unsound, loose = v[8 * kind + 2 * k], v[8 * kind + 2 * k + 1]
            assert unsound == 0, ("water" if kind == 0 else "wind", nm, unsound)
            assert loose == 0, ("water" if kind == 0 else "wind", nm, "conservative in", loose, "cases")


def test_out_of_order_execution_with_too_small_footprints_is_caught():
    """negative control of the test above: with every square of the other particle's sets shrunk by one ring the same
    adversarial order must NOT reproduce the reference - i.e. the test can see an unsound rule."""
    g = _golden.load("frame_rocksand_56")
    b = Backend(g)
    b.hs.lib.hs_set_mode(3, 0)
    b.hs.lib.hs_adversarial_weaken(1)
    try:
        b.hs.set_columns(_golden.cols(g, "init"))
        with pytest.raises(AssertionError):
            _golden.replay_frame(g, b, stats5)
    finally:
        b.hs.lib.hs_adversarial_weaken(0)
        b.hs.lib.hs_set_mode(0, 0)


@pytest.mark.parametrize("case", _golden.FRAME_CASES)
def test_exact_footprints_contain_every_access_of_a_step(case):
    """The exact-footprint schedule (sweep_exact, Foot<KIND> in sm_engine.cu) lets two steps overlap unless their
    footprints can meet: move() reads plus(ipos); a water step touches plus(ipos) U 3x3(npos) and writes
    {ipos} U 3x3(npos); a wind step touches and writes 5x5(ipos) U 5x5(npos).  Every record access of the cooperative
    step goes through one function of its backing store, so the host build records them by phase on the golden frames
    (split and staged exactly as the schedule does) and checks the sets really contain them - and that npos stays within
    the STEP the conflict range assumes (2 water, 3 wind)."""
    import ctypes as C
    g = _golden.load(case)
    b = Backend(g)
    out = (C.c_longlong * 6)()
    b.hs.lib.hs_set_mode(2, 0)
    try:
        b.hs.lib.hs_footprint_audit(out)          # reset
        b.hs.set_columns(_golden.cols(g, "init"))
        _golden.replay_frame(g, b, stats5)
        b.hs.lib.hs_footprint_audit(out)
    finally:
        b.hs.lib.hs_set_mode(0, 0)
    steps, in_move, in_interact, in_writeback, max_step, out_of_box = list(out)
    assert steps > 1000
    assert (in_move, in_interact, in_writeback) == (0, 0, 0)
    assert out_of_box == 0      # the box ipos +- R a particle publishes in its bin (particle_reach) holds the whole step
    assert max_step <= 3


@pytest.mark.parametrize("case", _golden.FRAME_CASES)
@pytest.mark.parametrize("lane_order", [0, 1], ids=["lanes_up", "lanes_down"])
def test_warp_cooperative_step_replays_golden_frame(case, lane_order):
    """sm_coop.cuh - the step as the sweep kernel's warps execute it (lane-parallel gathers and cascade
    evaluation, single-lane commits, staged 3x3 windows with write-back) - with the lanes run as loops on the
    host.  Both lane orders must reproduce the reference: inside a phase no lane may depend on another."""
    g = _golden.load(case)
    b = Backend(g)
    b.hs.lib.hs_set_mode(1, lane_order)
    try:
        b.hs.set_columns(_golden.cols(g, "init"))
        _golden.replay_frame(g, b, stats5)
    finally:
        b.hs.lib.hs_set_mode(0, 0)


@pytest.mark.parametrize("case", ["frame_rocksand_56", "frame_rgps_64"])
def test_mass_budget_matches_port_and_closes(case):
    """the six per-particle accumulators of the cooperative step (eroded, deposited, cascade_net,
    discarded, clamped, wind_negative) against the oracle port's, bit for bit, on a water and a wind batch; and the
    identity d(sum of heights) = deposited - eroded + cascade_net, to rounding."""
    from oracle import portapi
    g = _golden.load(case)
    b = Backend(g)
    b.hs.lib.hs_set_mode(1, 0)
    try:
        c0 = _golden.cols(g, "init")
        b.hs.set_columns(c0)
        po = portapi.Port().init(int(g["dimx"]), int(g["dimy"]), int(g["scale"]), g["soils"])
        po.set_columns(c0)
        for kind in ("water", "wind"):
            xy = g[kind + "_xy"]
            h0 = b.hs.heights().sum()
            (po.water_run if kind == "water" else po.wind_run)(xy)
            (b.water_run if kind == "water" else b.wind_run)(xy)
            per_p, sums_p = po.budget()
            per_h = b.hs.budget()
            _golden.same(per_h, per_p, kind + " budget per particle")
            dh = b.hs.heights().sum() - h0
            s = per_h.sum(axis=0)
            assert abs(dh - (s[1] - s[0] + s[2])) < 1e-9 * max(1.0, abs(s[0]) + abs(s[1])), (kind, dh, s)
            assert s[0] > 0 or s[1] > 0
            if kind == "water":
                assert s[3] > 0           # evaporating particles take sediment with them
    finally:
        b.hs.lib.hs_set_mode(0, 0)


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


# ---- pooling hydrology: the product's frame machine and active-cell index (sm_hydro.cuh) ----------------
class HydroBackend(Backend):
    def __init__(self, g, seep_mode):
        super().__init__(g)
        self.seep_mode = seep_mode

    def water_flood(self):
        return self.hs.water_flood()

    def seep(self):
        return self.hs.seep(self.seep_mode)


@pytest.mark.parametrize("case", _golden.HYDRO_CASES)
@pytest.mark.parametrize("seep_mode", [0, 1], ids=["every_cell", "active_index"])
def test_product_hydrology_replays_golden(case, seep_mode):
    """flood / water-table cascade / seep as the device executes them - explicit frame stack instead of
    recursion, and (seep_mode 1) the classification + active-cell index instead of the full scan - against
    vectors generated from the reference's recursive code."""
    g = _golden.load(case)
    b = HydroBackend(g, seep_mode)
    b.hs.set_columns(_golden.cols(g, "init"))
    counters = _golden.replay_hydro(g, b)
    assert all(c.overflow == 0 for c in counters)
    assert all(c.floods >= f for c, f in zip(counters, g["floods"]))


@pytest.mark.parametrize("case", _golden.HYDRO_CASES)
@pytest.mark.parametrize("seep_mode", [0, 1], ids=["every_cell", "active_index"])
@pytest.mark.parametrize("lane_order", [0, 1], ids=["lanes_up", "lanes_down"])
def test_warp_cooperative_hydrology_replays_golden(case, seep_mode, lane_order):
    """sm_hydro_coop.cuh - floods, water-table cascade frames evaluated eight neighbours at a time, nested
    particles on the cooperative step, the seep pass - with the lanes run as loops on the host, against the vectors
    generated from the reference's recursive code; the water batches themselves run the cooperative step too."""
    g = _golden.load(case)
    b = HydroBackend(g, seep_mode)
    b.hs.lib.hs_set_mode(1, lane_order)
    try:
        b.hs.set_columns(_golden.cols(g, "init"))
        counters = _golden.replay_hydro(g, b)
        assert all(c.overflow == 0 for c in counters)
        assert all(c.floods >= f for c, f in zip(counters, g["floods"]))
    finally:
        b.hs.lib.hs_set_mode(0, 0)


@pytest.mark.parametrize("coop", [0, 1], ids=["thread", "warp"])
def test_product_hydrology_counters_match_port(coop):
    """nested particles, their steps and the transfers are invisible to the verbatim reference; the oracle
    port (recursive, pinned to the reference) and the product's frame machine - one thread (sm_hydro.cuh) or one
    warp (sm_hydro_coop.cuh) - must agree on them."""
    from oracle import portapi
    g = _golden.load("hydro_bigbutte_40")
    b = HydroBackend(g, 1)
    b.hs.lib.hs_set_mode(coop, 0)
    b.hs.set_columns(_golden.cols(g, "init"))
    po = portapi.Port().init(int(g["dimx"]), int(g["dimy"]), int(g["scale"]), g["soils"])
    po.set_columns(_golden.cols(g, "init"))
    cells = int(g["dimx"]) * int(g["dimy"])
    for f in range(int(g["frames"])):
        xy = g["water_xy_%d" % f]
        po.water_run(xy); b.water_run(xy)
        x, y = po.water_flood(), b.water_flood()
        assert (x.floods, x.nested, x.nested_steps, x.transfers) == (y.floods, y.nested, y.nested_steps, y.transfers)
        x, y = po.seep(), b.seep()
        assert (x.floods, x.nested, x.nested_steps, x.transfers) == (y.floods, y.nested, y.nested_steps, y.transfers)
        assert x.cells == cells and y.cells < cells
        po.frequency_update(); b.frequency_update()
    b.hs.lib.hs_set_mode(0, 0)


def test_active_index_next_and_set():
    """the hierarchical bitmap of the seep pass (active_set / active_next) against a sorted list"""
    import ctypes as C
    hs = _hostsim.HostSim()
    rng = np.random.RandomState(3)
    for cells in (1, 63, 64, 65, 4096, 4097, 300000):
        k = min(cells, 200)
        idx = np.unique(rng.randint(0, cells, k)).astype(np.uint64)
        if cells > 64:
            idx = np.unique(np.concatenate([idx, np.array([0, cells - 1, 63, 64], np.uint64)]))
        out = np.zeros(len(idx) + 1, np.uint64)
        n = hs.lib.hs_active_selftest(C.c_uint64(cells), idx.ctypes.data_as(C.c_void_p), len(idx),
                                      out.ctypes.data_as(C.c_void_p))
        assert n == len(idx)
        assert np.array_equal(out[:n], idx)
