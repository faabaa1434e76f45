"""CPU: the oracle restatement (oracle/sm_oracle.cpp) against the golden vectors generated from the
reference itself, and - where oracle/_ref is built - against the reference directly."""
import numpy as np
import pytest
import _golden
from oracle import portapi, refapi


def stats5(st):
    return (st.steps, st.sweeps, st.exit_oob, st.exit_evap, st.exit_stall)


@pytest.fixture(scope="module")
def port():
    return portapi.Port()


@pytest.mark.parametrize("case", _golden.FRAME_CASES)
def test_port_replays_golden_frame(port, case):
    g = _golden.load(case)
    port.init(int(g["dimx"]), int(g["dimy"]), int(g["scale"]), g["soils"])
    port.set_columns(_golden.cols(g, "init"))
    _golden.same_cols(port.columns(), _golden.cols(g, "init"), "init columns")
    _golden.replay_frame(g, port, stats5)


def test_port_column_ops_truth_table(port):
    g = _golden.load("column_ops")
    port.init(8, 8, int(g["scale"]), g["soils"])
    for (kind, x, y, v, t), want in zip(g["ops"], g["remove_results"]):
        if kind == 0:
            port.add(int(x), int(y), v, int(t))
        else:
            got = port.remove(int(x), int(y), v)
            assert np.float64(got).tobytes() == np.float64(want).tobytes()
    _golden.same_cols(port.columns(), _golden.cols(g, "final"), "columns after add/remove")
    _golden.same(port.heights(), g["heights"], "heights")
    normals = np.array([port.normal(x, y) for x in range(8) for y in range(8)], np.float32)
    _golden.same(normals, g["normals"], "normals")
    bil = np.array([port.height(float(p[0]), float(p[1])) for p in g["bilinear_pts"]])
    _golden.same(bil, g["bilinear"], "bilinear heights")
    for x, y, loop in g["cascades"]:
        port.cascade(x, y, int(loop))
    _golden.same_cols(port.columns(), _golden.cols(g, "after_cascade"), "columns after cascades")


def test_port_edge_cases(port):
    """empty map, empty batches, particles spawned on the border / on empty columns."""
    g = _golden.load("column_ops")
    port.init(8, 8, int(g["scale"]), g["soils"])
    st = port.water_run(np.zeros((0, 2), np.float32))
    assert stats5(st) == (0, 0, 0, 0, 0)
    xy = np.array([[0, 0], [7, 7], [3, 4], [0, 7]], np.float32)   # all-empty map: normal is NaN -> out of bounds
    st = port.water_run(xy)
    assert st.steps == 0 and st.exit_oob + st.exit_stall == 4
    assert port.lib.smo_nsections() == 0
    assert port.remove(2, 2, 0.5) == 0.0                           # remove on an empty column


@pytest.mark.parametrize("soil,dim,n", [("rocksand", 72, 300), ("rockgravelpebblessand", 80, 300)])
def test_port_matches_reference_live(port, ref, soil, dim, n):
    ref.init(soil, seed=11, dimx=dim, dimy=dim - 8)
    port.init(ref.dimx, ref.dimy, ref.scale, ref.soils())
    port.set_columns(ref.columns())
    xw = ref.spawn_list(n, seed=11); xd = ref.spawn_list(n // 2)
    a, b = ref.water_run(xw), port.water_run(xw)
    assert stats5(a) == stats5(b)
    a, b = ref.wind_run(xd), port.wind_run(xd)
    assert (a.steps, a.exit_oob) == (b.steps, b.exit_oob)
    _golden.same_cols(ref.columns(), port.columns(), "columns")
    # sequential mode (the reference's own loop order) as well
    ref.init(soil, seed=11, dimx=dim, dimy=dim - 8)
    port.init(ref.dimx, ref.dimy, ref.scale, ref.soils()); port.set_columns(ref.columns())
    a, b = ref.water_seq(0, xw), port.water_seq(xw)
    assert stats5(a)[0] == stats5(b)[0]
    a, b = ref.wind_seq(0, xd), port.wind_seq(xd)
    assert a.steps == b.steps
    _golden.same_cols(ref.columns(), port.columns(), "columns (sequential)")


def test_port_vs_reference_random_column_ops(port, ref):
    """5000 random add / remove / cascade calls (incl. Air sections, zero and negative sizes, re-cascade
    budgets 0..3, border cells) on a small map: the restatement must track the reference section by section."""
    ref.init("rockgravelpebblessand", seed=2, dimx=12, dimy=10)
    port.init(ref.dimx, ref.dimy, ref.scale, ref.soils())
    port.set_columns(ref.columns())
    rng = np.random.RandomState(11)
    for i in range(5000):
        x, y = int(rng.randint(0, 12)), int(rng.randint(0, 10))
        k = rng.randint(0, 3)
        if k == 0:
            size = float(rng.choice([0.02, 0.0, -0.01, 0.3, 1e-9]) * rng.rand())
            typ = int(rng.randint(0, 5))
            ref.add(x, y, size, typ); port.add(x, y, size, typ)
        elif k == 1:
            h = float(rng.choice([0.02, 0.0, -0.5, 0.6, 1e-10]) * rng.rand())
            a, b = ref.remove(x, y, h), port.remove(x, y, h)
            assert np.float64(a).tobytes() == np.float64(b).tobytes(), i
        else:
            fx, fy, loop = float(rng.rand() * 11), float(rng.rand() * 9), int(rng.randint(0, 4))
            ref.cascade(fx, fy, loop); port.cascade(fx, fy, loop)
        if i % 500 == 499:
            _golden.same_cols(ref.columns(), port.columns(), "columns after %d ops" % (i + 1))
    _golden.same(ref.heights(), port.heights(), "heights")


# ---- pooling hydrology (flood, water-table cascade, seep) ------------------
@pytest.mark.parametrize("case", _golden.HYDRO_CASES)
def test_port_replays_golden_hydrology(port, case):
    g = _golden.load(case)
    port.init(int(g["dimx"]), int(g["dimy"]), int(g["scale"]), g["soils"])
    port.set_columns(_golden.cols(g, "init"))
    counters = _golden.replay_hydro(g, port)
    # the batch's own floods are what the reference can count; nested particles flood on top of that
    assert all(c.floods >= f for c, f in zip(counters, g["floods"]))
    assert sum(c.floods for c in counters) > 0 and sum(c.transfers for c in counters) > 0


@pytest.mark.parametrize("soil,dim,n,frames", [("default", 96, 500, 3), ("rockgravelpebblessand", 112, 900, 3),
                                               ("bigbutte", 64, 500, 3)])
def test_port_hydrology_matches_reference_live(port, ref, soil, dim, n, frames):
    """batch order: lockstep batch, floods in ascending index, seep pass - after every phase."""
    ref.init(soil, seed=42, dimx=dim, dimy=dim + 8)
    port.init(ref.dimx, ref.dimy, ref.scale, ref.soils())
    port.set_columns(ref.columns())
    for f in range(frames):
        xy = ref.spawn_list(n, seed=42 + f)
        ref.water_run(xy); port.water_run(xy)
        nf = ref.water_flood(); h = port.water_flood()
        assert h.floods >= nf
        _golden.same_cols(ref.columns(), port.columns(), "frame %d after floods" % f)
        ref.seep(); port.seep()
        _golden.same_cols(ref.columns(), port.columns(), "frame %d after seep" % f)
        ref.frequency_update(); port.frequency_update()
        fa, fb = ref.frequency(), port.frequency()
        for k in fa:
            _golden.same(fa[k], fb[k], k)


def test_port_hydrology_sequential_order(port, ref):
    """The reference's own order - each particle floods right after its loop, then the seep pass
    (SoilMachine.cpp:288-301) - with an explicit spawn list on both sides."""
    ref.init("default", seed=42, dimx=96, dimy=96)
    port.init(ref.dimx, ref.dimy, ref.scale, ref.soils())
    port.set_columns(ref.columns())
    for f in range(2):
        xy = ref.spawn_list(400, seed=3 + f)
        a = ref.water_seq(0, xy, flood=True, seep=True)
        b, h = port.water_seq_full(xy, flood=True, seep=True)
        assert stats5(a)[0] == stats5(b)[0] and (a.exit_oob, a.exit_evap, a.exit_stall) == (b.exit_oob, b.exit_evap, b.exit_stall)
        _golden.same_cols(ref.columns(), port.columns(), "frame %d" % f)
        ref.frequency_update(); port.frequency_update()
    assert h.cells == ref.dimx * ref.dimy


def test_reference_build_reproduces_survey_hydrology_kat(ref):
    """default.soil, 256^2, SEED 42, srand(42), one frame of 1000 water particles through
    the reference loop INCLUDING flood and the seep pass (rand() spawns, sequential order)."""
    ref.init("default", seed=42, dimx=256, dimy=256)
    ref.lib.smref_srand(42)
    st = ref.water_seq(1000, None, flood=True, seep=True)
    c = ref.columns()
    assert st.steps == 313844 and (st.exit_oob, st.exit_evap, st.exit_stall) == (170, 636, 194)
    assert len(c["type"]) == 65638 and int((c["type"] == 0).sum()) == 309
    assert abs(ref.heights().sum() - 29250.672765019299) < 1e-9


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_hydrology_random_differential(port, ref, seed):
    """Random wet maps (0-4 soil sections per column with random saturations, Air on top of a quarter of
    them, empty columns, tiny sections), then 40 random operations - seep pass, small water batch + its
    floods, puddles added by hand, frequency update - on the reference, the oracle port and the product core
    (tests/hostsim, alternating between the every-cell and the active-index seep pass): all three must stay
    identical section by section, saturation included."""
    import _hostsim
    ref.init("rockgravelpebblessand", seed=3, dimx=14, dimy=11)
    soils = ref.soils()
    rng = np.random.RandomState(seed)
    off, typ, size, sat = [0], [], [], []
    for c in range(ref.dimx * ref.dimy):
        k = int(rng.randint(0, 5))
        for j in range(k):
            typ.append(int(rng.randint(1, len(soils))))
            size.append(float(rng.choice([0.05, 0.2, 1e-7, 0.01]) * rng.rand() + 0.001))
            sat.append(float(rng.choice([0, 0, 0.3, 1.0, 0.999])))
        if rng.rand() < 0.25:
            typ.append(0); size.append(float(0.05 * rng.rand() + 1e-4)); sat.append(float(rng.choice([0, 1.0])))
            k += 1
        off.append(off[-1] + k)
    ref.set_columns(np.array(off, np.int64), np.array(typ, np.int32), np.array(size), np.array(sat))
    cols = ref.columns()
    port.init(ref.dimx, ref.dimy, ref.scale, soils); port.set_columns(cols)
    hs = _hostsim.HostSim(); hs.init(ref.dimx, ref.dimy, ref.scale, soils); hs.set_columns(cols)

    def check(tag):
        a = ref.columns()
        _golden.same_cols(a, port.columns(), tag + " (port)")
        _golden.same_cols(a, hs.columns(), tag + " (product core)")

    check("initial")
    for it in range(40):
        op = int(rng.randint(0, 4))
        if op == 0:
            ref.seep(); port.seep(); hs.seep(int(rng.randint(0, 2)))
        elif op == 1:
            n = int(rng.randint(1, 30))
            xy = np.stack([rng.randint(0, ref.dimx, n), rng.randint(0, ref.dimy, n)], 1).astype(np.float32)
            ref.water_run(xy); port.water_run(xy); hs.water_run(xy)
            check("op %d batch" % it)
            ref.water_flood(); port.water_flood(); hs.water_flood()
        elif op == 2:
            for _ in range(5):
                x, y, s = int(rng.randint(0, ref.dimx)), int(rng.randint(0, ref.dimy)), float(0.2 * rng.rand())
                ref.add(x, y, s, 0); port.add(x, y, s, 0); hs.lib.hs_add(x, y, s, 0)
        else:
            ref.frequency_update(); port.frequency_update(); hs.frequency_update()
        check("op %d kind %d" % (it, op))


def test_hydrology_edge_cases(port, ref):
    """empty map, empty batch, puddles on the corners and edges of the map (the water cascade and the
    active-cell index both clip their 3x3 blocks), a lone puddle on an otherwise empty map."""
    import _hostsim
    ref.init("rocksand", seed=9, dimx=10, dimy=7)
    soils = ref.soils()
    hs = _hostsim.HostSim()

    def load(cols):
        port.init(ref.dimx, ref.dimy, ref.scale, soils); port.set_columns(cols)
        hs.init(ref.dimx, ref.dimy, ref.scale, soils); hs.set_columns(cols)

    def check(tag):
        a = ref.columns()
        _golden.same_cols(a, port.columns(), tag + " (port)")
        _golden.same_cols(a, hs.columns(), tag + " (product core)")

    # 1. all-empty map: nothing to visit, nothing to flood
    cells = ref.dimx * ref.dimy
    empty = np.zeros(cells + 1, np.int64)
    ref.set_columns(empty, np.zeros(0, np.int32), np.zeros(0))
    load(ref.columns())
    none = np.zeros((0, 2), np.float32)
    ref.water_run(none); port.water_run(none); hs.water_run(none)
    assert ref.water_flood() == 0 and port.water_flood().floods == 0 and hs.water_flood().floods == 0
    ref.seep(); port.seep()
    assert hs.seep(1).cells == 0                   # the active index finds no wet cell at all
    check("empty map")
    # 2. a lone puddle on the empty map: water spreads over bare ground, nested particles run on height 0
    ref.add(4, 3, 0.3, 0); port.add(4, 3, 0.3, 0); hs.lib.hs_add(4, 3, 0.3, 0)
    for k in range(3):
        ref.seep(); port.seep(); hs.seep(k % 2)
        check("lone puddle pass %d" % k)
    # 3. terrain with puddles on every corner and along the edges
    ref.init("rocksand", seed=9, dimx=10, dimy=7)
    load(ref.columns())
    spots = [(0, 0), (9, 0), (0, 6), (9, 6), (5, 0), (5, 6), (0, 3), (9, 3), (4, 4)]
    for x, y in spots:
        ref.add(x, y, 0.08, 0); port.add(x, y, 0.08, 0); hs.lib.hs_add(x, y, 0.08, 0)
    for k in range(4):
        ref.seep(); port.seep(); hs.seep((k + 1) % 2)
        check("border puddles pass %d" % k)
    xy = np.array([[0, 0], [9, 6], [0, 6], [9, 0], [5, 3], [8, 5]], np.float32)   # spawns on the corners too
    ref.water_run(xy); port.water_run(xy); hs.water_run(xy)
    ref.water_flood(); port.water_flood(); hs.water_flood()
    check("floods on the border")
