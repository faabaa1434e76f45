"""First-light probe run on the GPU box: timings of a few batches (not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refapi
import soilmachine_b200 as smb

def run(soil, dim, nw, nd, lanes=None):
    if lanes: os.environ["SM_LANES"] = str(lanes)
    r = refapi.get().init(soil, seed=42, dimx=dim, dimy=dim, poolsize=dim*dim*4+2000000)
    ctx = smb.Context(r.dimx, r.dimy, r.scale, max_particles=max(nw, nd, 1))
    ctx.set_soils(r.soils())
    cols = r.columns()
    ctx.upload_columns(cols["offsets"], cols["type"], cols["size"], cols["saturation"])
    r.lib.smref_srand(42)
    xw = r.spawn_list(nw); xd = r.spawn_list(nd)
    for it in range(2):
        t0=time.time(); g = ctx.water_run(xw); t1=time.time()
        print(soil, dim, "lanes", lanes, "water n=%d steps=%d sweeps=%d ms=%.2f -> %.3e steps/s (wall %.3f)" % (nw, g.steps, g.sweeps, g.device_ms, g.steps/g.device_ms*1e3, t1-t0), flush=True)
        if nd:
            g = ctx.wind_run(xd)
            print(soil, dim, "lanes", lanes, "wind  n=%d steps=%d sweeps=%d ms=%.2f -> %.3e steps/s" % (nd, g.steps, g.sweeps, g.device_ms, g.steps/max(g.device_ms,1e-9)*1e3), flush=True)
    ctx.close()

if __name__ == "__main__":
    run("rocksand", 1024, 10000, 10000)
    for lanes in (1, 4, 32):
        run("rocksand", 1024, 10000, 0, lanes)
