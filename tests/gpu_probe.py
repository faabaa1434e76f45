"""Timing probe run on the GPU box (not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refapi
import soilmachine_b200 as smb

def run(soil, dim, nw, nd, lanes=None, iters=1, label=""):
    if lanes: os.environ["SM_LANES"] = str(lanes)
    r = refapi.get().init(soil, seed=42, dimx=dim, dimy=dim, poolsize=dim*dim*4+2000000)
    ctx = smb.Context(r.dimx, r.dimy, r.scale, max_particles=max(nw, nd, 1))
    ctx.set_soils(r.soils())
    cols = r.columns()
    ctx.upload_columns(cols["offsets"], cols["type"], cols["size"], cols["saturation"])
    r.lib.smref_srand(42)
    xw = r.spawn_list(nw) if nw else None
    xd = r.spawn_list(nd) if nd else None
    for it in range(iters):
        if nw:
            g = ctx.water_run(xw)
            print(label, soil, dim, "lanes", lanes, "water n=%d steps=%d sweeps=%d ms=%.2f -> %.3e steps/s, %.2f us/sweep" % (nw, g.steps, g.sweeps, g.device_ms, g.steps/g.device_ms*1e3, g.device_ms*1e3/max(g.sweeps,1)), flush=True)
        if nd:
            g = ctx.wind_run(xd)
            print(label, soil, dim, "lanes", lanes, "wind  n=%d steps=%d sweeps=%d ms=%.2f -> %.3e steps/s, %.2f us/sweep" % (nd, g.steps, g.sweeps, g.device_ms, g.steps/max(g.device_ms,1e-9)*1e3, g.device_ms*1e3/max(g.sweeps,1)), flush=True)
    ctx.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("profile", "one", "mesh", "big", "tail", "poolrate", "sweepcurve", "sweepstat") or which.startswith("cfg3:"): which = "none"
    if which in ("all", "single"):
        # single particles: sweep time = step latency (+ trivial barrier)
        run("rocksand", 1024, 1, 0, label="single")
        run("rocksand", 1024, 0, 64, label="single")   # 64 wind particles (half die at once), sparse
        run("rocksand", 1024, 64, 0, label="sparse64")
    if which in ("all", "mid"):
        run("rocksand", 1024, 10000, 10000, label="mid")
    if which in ("all", "cfg3d"):
        run("rockgravelpebblessand", 1024, 1563, 1563, label="cfg3-density")


def profile(soil, dim, n, kind, lanes=None):
    """Phase breakdown (needs a -DSM_PROFILE build)."""
    import ctypes as C
    if lanes: os.environ["SM_LANES"] = str(lanes)
    r = refapi.get().init(soil, seed=42, dimx=dim, dimy=dim, poolsize=dim*dim*4+2000000)
    ctx = smb.Context(r.dimx, r.dimy, r.scale, max_particles=max(n, 1))
    ctx.set_soils(r.soils())
    cols = r.columns()
    ctx.upload_columns(cols["offsets"], cols["type"], cols["size"], cols["saturation"])
    r.lib.smref_srand(42)
    xy = r.spawn_list(n)
    out = (C.c_uint64 * 16)()
    ctx.lib.sm_debug_profile(ctx.h, out, 1)
    g = ctx.water_run(xy) if kind == "water" else ctx.wind_run(xy)
    ctx.lib.sm_debug_profile(ctx.h, out, 1)
    names = ["looptop", "stateload", "wait", "fence_acq", "step", "writeback", "fence_rel", "barrier",
             "  step.begin(fetchA)", "  step.move", "  step.fetchB", "  step.interact", "blocked(waiting)",
             "    interact.bilinear+c_eq", "    interact.erode/deposit", "    interact.cascade(+tail)"]
    tot = sum(out[i] for i in range(8)) + out[12]
    print("PROFILE", soil, dim, kind, "n=%d lanes=%s steps=%d sweeps=%d ms=%.2f us/sweep=%.2f" % (n, lanes, g.steps, g.sweeps, g.device_ms, g.device_ms*1e3/max(g.sweeps,1)))
    for i, nm in enumerate(names):
        print("   %-22s %8.0f cycles/step  %5.1f%%" % (nm, out[i] / max(g.steps, 1), 100.0 * out[i] / max(tot, 1)))
    ctx.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "profile":
    profile("rocksand", 1024, 1, "water")
    profile("rocksand", 1024, 64, "water")
    profile("rocksand", 1024, 64, "wind")
    profile("rockgravelpebblessand", 1024, 1563, "water")
    profile("rockgravelpebblessand", 1024, 1563, "wind")


def cfg3(kind="wind", frames=2):
    """config-3 batches through the product's own terrain init (no oracle)."""
    from soilmachine_b200 import host
    sim = host.Simulation("rockgravelpebblessand", seed=42, dimx=4096, dimy=4096, max_particles=25000)
    for f in range(frames):
        xw = host.spawn_list(25000, 4096, 4096); xd = host.spawn_list(25000, 4096, 4096)
        if kind in ("water", "both"):
            g = sim.ctx.water_run(xw)
            print("cfg3 water lanes=%s async=%s: steps=%d sweeps=%d ms=%.1f -> %.3e steps/s" % (os.environ.get("SM_LANES"), os.environ.get("SM_ASYNC"), g.steps, g.sweeps, g.device_ms, g.steps / g.device_ms * 1e3), flush=True)
        if kind in ("wind", "both"):
            g = sim.ctx.wind_run(xd)
            print("cfg3 wind  lanes=%s async=%s: steps=%d sweeps=%d ms=%.1f -> %.3e steps/s" % (os.environ.get("SM_LANES"), os.environ.get("SM_ASYNC"), g.steps, g.sweeps, g.device_ms, g.steps / g.device_ms * 1e3), flush=True)
    sim.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1].startswith("cfg3:"):
    cfg3(sys.argv[1].split(":")[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "one":
    run("rocksand", 1024, 1 if sys.argv[2] == "water" else 0, 0 if sys.argv[2] == "water" else 2, label="one")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "mesh":
    from soilmachine_b200 import host
    sim = host.Simulation("rockgravelpebblessand", seed=42, dimx=4096, dimy=4096, max_particles=1024)
    for it in range(5):
        sim.ctx.mesh_update(240, download=False)
        st = sim.ctx.last_stats()
        cells = 4096 * 4096
        print("k_mesh 4096^2: %.3f ms, %.1f GB/s algorithmic (32 B in + 44 B out per cell)" % (st.device_ms, cells * 76 / st.device_ms / 1e6), flush=True)
    sim.ctx.timer_start(); sim.ctx.frequency_update(); ms = sim.ctx.timer_stop()
    print("k_frequency_update 4096^2: %.3f ms, %.1f GB/s (16 B/cell)" % (ms, cells * 16 / ms / 1e6))


def bigcfg(soil, dim, nw, nd, frames=2):
    from soilmachine_b200 import host
    t0 = time.time()
    sim = host.Simulation(soil, seed=42, dimx=dim, dimy=dim, max_particles=max(nw, nd, 1))
    print("%s %d^2: terrain init + context %.2f s, %d sections" % (soil, dim, time.time() - t0, sim.ctx.section_count() if dim <= 4096 else -1), flush=True)
    for f in range(frames):
        xw = host.spawn_list(nw, dim, dim)
        g = sim.ctx.water_run(xw)
        print("  frame %d water n=%d: steps=%d sweeps=%d ms=%.1f -> %.3e steps/s drops=%d" % (f, nw, g.steps, g.sweeps, g.device_ms, g.steps / g.device_ms * 1e3, g.pool_drops), flush=True)
        if nd:
            xd = host.spawn_list(nd, dim, dim)
            g = sim.ctx.wind_run(xd)
            print("  frame %d wind  n=%d: steps=%d sweeps=%d ms=%.1f exits=%d" % (f, nd, g.steps, g.sweeps, g.device_ms, g.exit_oob), flush=True)
        sim.ctx.frequency_update()
    print("  height sum %.9f" % sim.ctx.height_sum(), flush=True)
    sim.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "big":
    bigcfg("bigbutte", 4096, 50000, 0)                       # BASELINE config 4 shape (single GPU)
    bigcfg("rockgravelpebbles_big", 8192, 100000, 100000)    # BASELINE config 5 shape: 200k mixed, wind inert


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "tail":
    import ctypes as C
    from soilmachine_b200 import host
    sim = host.Simulation("rockgravelpebblessand", seed=42, dimx=4096, dimy=4096, max_particles=25000)
    out = (C.c_uint64 * 16)()
    for f in range(2):
        xw = host.spawn_list(25000, 4096, 4096); xd = host.spawn_list(25000, 4096, 4096)
        sim.ctx.water_run(xw)
        sim.ctx.lib.sm_debug_profile(sim.ctx.h, out, 1)
        g = sim.ctx.wind_run(xd)
        sim.ctx.lib.sm_debug_profile(sim.ctx.h, out, 2)
        print("wind frame %d: steps=%d sweeps=%d ms=%.1f | steps>20k cycles: %d (%.2f/sweep), >40k: %d (%.2f/sweep), max step %d cycles; avg step %.0f cycles"
              % (f, g.steps, g.sweeps, g.device_ms, out[13], out[13] / g.sweeps, out[14], out[14] / g.sweeps, out[15], out[4] / max(g.steps, 1)), flush=True)
        print("   transfers per step: %.3f overall; in the >40k-cycle steps: %.2f transfers, %.0f cycles on average"
              % (out[10] / max(g.steps, 1), out[11] / max(out[14], 1), out[12] / max(out[14], 1)), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "poolrate":
    import ctypes as C
    from soilmachine_b200 import host
    sim = host.Simulation("rockgravelpebblessand", seed=42, dimx=4096, dimy=4096, max_particles=25000)
    out = (C.c_uint64 * 16)()
    def snap():
        sim.ctx.lib.sm_debug_profile(sim.ctx.h, out, 3)
        return out[0] + out[1], out[2] + out[3], out[4]
    for f in range(2):
        xw = host.spawn_list(25000, 4096, 4096); xd = host.spawn_list(25000, 4096, 4096)
        a0 = snap(); g = sim.ctx.water_run(xw); a1 = snap()
        print("water: steps=%d frees=%d (%.3f/step) ring-allocs=%d bump-allocs=%d (allocs %.3f/step)" % (g.steps, a1[0]-a0[0], (a1[0]-a0[0])/g.steps, a1[1]-a0[1], a1[2]-a0[2], (a1[1]-a0[1]+a1[2]-a0[2])/g.steps), flush=True)
        g = sim.ctx.wind_run(xd); a2 = snap()
        print("wind : steps=%d frees=%d (%.3f/step) ring-allocs=%d bump-allocs=%d (allocs %.3f/step)" % (g.steps, a2[0]-a1[0], (a2[0]-a1[0])/g.steps, a2[1]-a1[1], a2[2]-a1[2], (a2[1]-a1[1]+a2[2]-a1[2])/g.steps), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sweepcurve":
    import ctypes as C
    from soilmachine_b200 import host
    sim = host.Simulation("rockgravelpebblessand", seed=42, dimx=4096, dimy=4096, max_particles=25000)
    xw = host.spawn_list(25000, 4096, 4096); xd = host.spawn_list(25000, 4096, 4096)
    sim.ctx.water_run(xw)
    g = sim.ctx.wind_run(xd)
    n = min(int(g.sweeps), 16384)
    buf = np.zeros((n, 2), np.uint64)
    sim.ctx.lib.sm_debug_sweeps(sim.ctx.h, buf.ctypes.data_as(C.c_void_p), n)
    t = buf[:, 0].astype(np.float64); alive = buf[:, 1].astype(np.int64)
    dt = np.diff(t) / 1.965e3    # us at 1965 MHz
    print("wind: sweeps=%d ms=%.1f" % (g.sweeps, g.device_ms))
    for lo in range(0, n - 1, 1000):
        hi = min(lo + 1000, n - 1)
        print("  sweeps %5d-%5d: alive %6d..%6d  avg %.1f us/sweep  (sum %.1f ms)" % (lo, hi, alive[lo], alive[hi], dt[lo:hi].mean(), dt[lo:hi].sum() / 1e3), flush=True)


def hydro(soil="rockgravelpebblessand", dim=4096, n=25000, frames=4):
    """full water part of the frame at bench scale: batch, floods, seep pass (device timings)."""
    from soilmachine_b200 import host
    sim = host.Simulation(soil, seed=42, dimx=dim, dimy=dim, max_particles=n)
    for f in range(frames):
        xw = host.spawn_list(n, dim, dim)
        g = sim.ctx.water_run(xw)
        h = sim.ctx.water_flood()
        s = sim.ctx.seep()
        sim.ctx.frequency_update()
        print("hydro %s %d^2 frame %d: batch %.1f ms (%d steps, %d stalled) | flood %.2f ms %s | seep %.2f ms %s" % (
            soil, dim, f, g.device_ms, g.steps, g.exit_stall, h.device_ms,
            {k: v for k, v in h.asdict().items() if k != "device_ms"}, s.device_ms,
            {k: v for k, v in s.asdict().items() if k != "device_ms"}), flush=True)
    print("  height sum %.9f" % sim.ctx.height_sum(), flush=True)
    sim.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1].startswith("hydro"):
    hydro()
    if sys.argv[1] == "hydro":
        hydro("default", 1024, 10000, 4)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sweepstat":
    # needs a -DSM_PROFILE build (SM_LIB_PATH): per-sweep statistics of the warp kernel's wind batch at config 3
    import ctypes as C
    from soilmachine_b200 import host
    sim = host.Simulation("rockgravelpebblessand", seed=42, dimx=4096, dimy=4096, max_particles=25000)
    xw = host.spawn_list(25000, 4096, 4096); xd = host.spawn_list(25000, 4096, 4096)
    for kind, xy in (("water", xw), ("wind", xd)):
        buf = np.zeros((16384, 8), np.uint64)
        sim.ctx.lib.sm_debug_sweeps8(sim.ctx.h, buf.ctypes.data_as(C.c_void_p), 16384)      # clear
        g = sim.ctx.water_run(xy) if kind == "water" else sim.ctx.wind_run(xy)
        n = min(int(g.sweeps), 16380)
        sim.ctx.lib.sm_debug_sweeps8(sim.ctx.h, buf.ctypes.data_as(C.c_void_p), 16384)
        ph = buf[16383].astype(np.float64)
        if ph[5] > 0:
            k = 1.0 / ph[5] / 1.965e3
            print("%s phases (us per step, conservative path, %d steps): load+scan %.2f | wait %.2f | fetch+move %.2f | interact %.2f | flush+publish %.2f (with a release fence: %.2f over %d steps)"
                  % (kind, ph[5], ph[0] * k, ph[1] * k, ph[2] * k, ph[3] * k, ph[4] * k, ph[6] / max(ph[7], 1) / 1.965e3, ph[7]))
        b = buf[:n].astype(np.float64)
        live, mxstep, mxwait, mxwarp, sumstep, nstep = b[:, 0], b[:, 1], b[:, 2], b[:, 3], b[:, 4], b[:, 5]
        t0 = (~buf[:n, 6]).astype(np.float64); t1 = b[:, 7]
        span = (t1 - t0) / 1e3                                  # us between the first warp's start and the last warp's end
        period = np.diff(t0) / 1e3                              # us from sweep to sweep
        cyc = 1.965e3
        print("%s: sweeps=%d ms=%.1f  us/sweep=%.2f" % (kind, g.sweeps, g.device_ms, g.device_ms * 1e3 / g.sweeps))
        print("  avg step %.2f us, avg of per-sweep max step %.2f us, avg max wait %.2f us, avg max warp busy %.2f us, avg span %.2f us, avg period %.2f us"
              % (sumstep.sum() / nstep.sum() / cyc, mxstep.mean() / cyc, mxwait.mean() / cyc, mxwarp.mean() / cyc, span.mean(), period.mean()))
        for lo in range(0, n - 1, max(n // 14, 1)):
            hi = min(lo + max(n // 14, 1), n - 1)
            sl = slice(lo, hi)
            print("  sweeps %5d-%5d live %6d: period %.1f us | span %.1f | max warp busy %.1f | max step %.1f | avg step %.2f | max wait %.1f | trips %.2f"
                  % (lo, hi, live[sl].mean(), period[sl].mean(), span[sl].mean(), mxwarp[sl].mean() / cyc, mxstep[sl].mean() / cyc,
                     sumstep[sl].sum() / max(nstep[sl].sum(), 1) / cyc, mxwait[sl].mean() / cyc, np.ceil(live[sl] / 3552).mean()), flush=True)
