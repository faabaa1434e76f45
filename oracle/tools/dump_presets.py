"""Regenerates soilmachine_b200/presets/*.json from the reference's own loader.

Run here (where /root/reference is mounted):  python oracle/tools/dump_presets.py
Each preset is what loadsoil() (source/io.h:7-230) leaves in the reference's tables for one
soil/*.soil file - including the loader's quirks (fields inherited from the previous SOIL block,
ids in order of first mention, placeholder entries) - dumped through oracle/_ref.  float32 values
are written as the exact decimal of the float, so they round-trip bit for bit.
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refapi  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "soilmachine_b200", "presets")
    os.makedirs(out_dir, exist_ok=True)
    r = refapi.get()
    src = "/root/reference/soil" if os.path.isdir("/root/reference/soil") else refapi.SOIL_DIR
    for path in sorted(glob.glob(os.path.join(src, "*.soil"))):
        name = os.path.basename(path)[:-5]
        r.init(path, seed=0, dimx=8, dimy=8, poolsize=1000)   # tiny map: only the tables matter
        soils, layers = r.soils(), r.layers()
        w = _world(r, path)
        doc = {
            "source": "soil/%s.soil via source/io.h:loadsoil" % name,
            "world": w,
            "soils": [dict(name=s["name"].decode(), **{k: (int(s[k]) if k in ("transports", "erodes", "cascades", "abrades")
                                                           else float(s[k])) for k in soils.dtype.names
                                                       if k not in ("name", "color")},
                           color=[float(c) for c in s["color"]]) for s in soils],
            "layers": [{k: (int(l[k]) if k == "type" else float(l[k])) for k in layers.dtype.names} for l in layers],
        }
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(doc, f, indent=1)
        print("wrote", name, len(soils), "soils", len(layers), "layers", w)


def _world(r, path):
    # WORLD block values (io.h:208-220); parsed here only to record the file's own defaults
    w = {"sizex": 256, "sizey": 256, "scale": 80, "nwater": 250, "nwind": 250}
    inside = False
    for line in open(path):
        line = line.split("#")[0].rstrip("\n")
        if line.startswith("WORLD"):
            inside = True
            continue
        if line == "}":
            inside = False
        if inside and " " in line:
            tag, val = line.split(" ", 1)
            if tag.lower() in w:
                w[tag.lower()] = int(val)
    assert w["scale"] == r.scale, (w, r.scale)
    return w


if __name__ == "__main__":
    main()
