"""Helpers shared by the golden-vector tests: replay a golden case on any backend that offers the
driver interface of oracle/refapi.py (init/set_columns/water_run/...)."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_CASES = ["frame_default_48", "frame_rocksand_56", "frame_rgps_64", "frame_bigbutte_40"]
HYDRO_CASES = ["hydro_default_48", "hydro_bigbutte_40"]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def cols(g, prefix):
    return {k: g[prefix + "_" + k] for k in ("offsets", "type", "size", "floor", "saturation")}


def same(a, b, what):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
        bad = np.nonzero(a.reshape(-1) != b.reshape(-1))[0]
        raise AssertionError("%s differs at %d entries, first %s: %r vs %r" %
                             (what, len(bad), bad[:4], a.reshape(-1)[bad[:4]], b.reshape(-1)[bad[:4]]))


def same_cols(c1, c2, what):
    for k in ("offsets", "type", "size", "floor", "saturation"):
        same(c1[k], c2[k], what + "." + k)


def replay_frame(g, backend, stats_of):
    """backend: object already initialised with g's map (init columns loaded).  stats_of(st) ->
    (steps, sweeps, oob, evap, stall)."""
    sw = backend.water_run(g["water_xy"])
    same(np.array(stats_of(sw), np.int64), g["water_stats"], "water stats")
    ws = backend.water_state()
    for k in ws:
        same(ws[k], g["water_state_" + k], "water state " + k)
    same_cols(backend_columns(backend), cols(g, "after_water"), "columns after water")
    if len(g["wind_xy"]):
        sd = backend.wind_run(g["wind_xy"])
        got = np.array(stats_of(sd), np.int64)
        same(got[[0, 2]], g["wind_stats"][[0, 2]], "wind stats (steps, exits)")
        ds = backend.wind_state()
        for k in ds:
            same(ds[k], g["wind_state_" + k], "wind state " + k)
    backend.frequency_update()
    same_cols(backend_columns(backend), cols(g, "after_frame"), "columns after frame")
    f = backend.frequency()
    for k in f:
        same(f[k], g["freq_" + k], "frequency " + k)
    same(backend.heights(), g["heights"], "heights")


def replay_hydro(g, backend):
    """backend initialised with g's init columns; needs water_run / water_flood / seep / frequency_update.
    Returns the per-frame counters of the backend's flood phase."""
    counters = []
    for f in range(int(g["frames"])):
        backend.water_run(g["water_xy_%d" % f])
        counters.append(backend.water_flood())
        if ("after_flood_%d_offsets" % f) in g.files:
            same_cols(backend_columns(backend), cols(g, "after_flood_%d" % f), "frame %d columns after the floods" % f)
        backend.seep()
        same_cols(backend_columns(backend), cols(g, "after_seep_%d" % f), "frame %d columns after the seep pass" % f)
        backend.frequency_update()
    fr = backend.frequency()
    for k in fr:
        same(fr[k], g["freq_" + k], "frequency " + k)
    same(backend.heights(), g["heights"], "heights")
    return counters


def backend_columns(b):
    return b.columns() if hasattr(b, "columns") else b.download_columns()
