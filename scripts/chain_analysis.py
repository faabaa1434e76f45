"""One-off analysis (CPU, oracle only): how long are the dependency chains of a water sweep late in a config-3
batch under (a) the conservative box rule, (b) exact footprints, and how crowded are the clusters."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refapi, portapi
from soilmachine_b200 import host

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
soil = sys.argv[3] if len(sys.argv) > 3 else "rockgravelpebblessand"
sweeps = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "50,200,400").split(",")]
t = time.time()
ref = refapi.get(); ref.init(soil, seed=42, dimx=dim, dimy=dim, poolsize=dim * dim * 2 + 2000000)
cols = ref.columns(); soils = ref.soils()
print("terrain %.1fs" % (time.time() - t), flush=True)
host.srand(42) if hasattr(host, "srand") else None
xw = host.spawn_list(n, dim, dim)

def state_after(k):
    p = portapi.Port(); p.init(dim, dim, ref.scale, soils); p.set_columns(cols)
    p.water_run(xw, max_sweeps=k)
    s = p.water_state()
    return np.rint(s["pos"]).astype(np.int64), s["alive"].astype(bool)

def depth(ip, np_, alive, rule):
    idx = np.nonzero(alive)[0]
    order = idx                                    # ascending index
    G = 8
    bins = {}
    dep = np.zeros(len(alive), np.int32); nb = np.zeros(len(alive), np.int32)
    for a in order:
        ax, ay = ip[a]; bx, by = np_[a]
        d = 0; cnt = 0
        for gx in range(ax // G - 1, ax // G + 2):
            for gy in range(ay // G - 1, ay // G + 2):
                for b in bins.get((gx, gy), ()):
                    cx, cy = ip[b]; dx, dy = np_[b]
                    if rule == "box":
                        hit = abs(ax - cx) <= 6 and abs(ay - cy) <= 6
                    else:
                        # F = plus(ipos) U 3x3(npos); two footprints meet if any pair of their parts does
                        def meets(p1, r1, p2, r2): return abs(p1[0] - p2[0]) <= r1 + r2 and abs(p1[1] - p2[1]) <= r1 + r2
                        hit = meets((ax, ay), 1, (cx, cy), 1) or meets((ax, ay), 1, (dx, dy), 1) or meets((bx, by), 1, (cx, cy), 1) or meets((bx, by), 1, (dx, dy), 1)
                    if hit:
                        cnt += 1
                        if dep[b] + 1 > d: d = dep[b] + 1
        dep[a] = d; nb[a] = cnt
        bins.setdefault((ax // G, ay // G), []).append(a)
    return dep[idx], nb[idx]

for k in sweeps:
    t = time.time()
    ip, al = state_after(k)
    np_, al2 = state_after(k + 1)
    print("sweep %d: alive %d (port %.1fs)" % (k, al.sum(), time.time() - t), flush=True)
    cells = {}
    for a in np.nonzero(al)[0]: cells[tuple(ip[a])] = cells.get(tuple(ip[a]), 0) + 1
    occ = np.array(sorted(cells.values(), reverse=True))
    print("   distinct cells %d, most crowded cells %s, particles in cells shared by >=2: %d" % (len(occ), occ[:8].tolist(), occ[occ >= 2].sum()))
    for rule in ("box", "exact"):
        dep, nb = depth(ip, np_, al, rule)
        print("   %-5s: longest chain %d, mean depth %.2f, lower-index neighbours: mean %.1f max %d, >31: %d particles"
              % (rule, dep.max() + 1, dep.mean(), nb.mean(), nb.max(), (nb > 31).sum()), flush=True)
