"""Extract the reference's binary test fixtures into portable files (run in the build
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py

Sources (read-only): /root/reference/test/inputs/{test_mc_points,stress_strain,floe_shapes}.jld2
Outputs (committed):  tests/golden/{test_mc_points,stress_strain}.json, floe_shapes.npz
The numeric golden VALUES asserted by the reference's tests are transcribed in
tests/golden/reference_values.py with their file:line.
"""
import json
import os
import numpy as np
from jld2_mini import JLD2File

SRC = "/root/reference/test/inputs"
OUT = os.path.dirname(os.path.abspath(__file__))


def tolist(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, list):
        return [tolist(x) for x in v]
    return float(v)


def main():
    f = JLD2File(os.path.join(SRC, "test_mc_points.jld2"))
    json.dump({"X": tolist(f["X"]), "Y": tolist(f["Y"])},
              open(os.path.join(OUT, "test_mc_points.json"), "w"))
    f = JLD2File(os.path.join(SRC, "stress_strain.jld2"))
    json.dump({k: tolist(f[k]) for k in f.keys()},
              open(os.path.join(OUT, "stress_strain.json"), "w"))
    f = JLD2File(os.path.join(SRC, "floe_shapes.jld2"))
    polys = f["floe_vertices"]
    offs, xy = [0], []
    for p in polys:
        ring = p[0] if isinstance(p, list) and isinstance(p[0], list) else p
        pts = np.array([np.asarray(q, dtype=np.float64) for q in ring])
        xy.append(pts)
        offs.append(offs[-1] + len(pts))
    np.savez_compressed(os.path.join(OUT, "floe_shapes.npz"),
                        offsets=np.array(offs, dtype=np.int64), xy=np.concatenate(xy))
    print("polys", len(polys), "verts", offs[-1])


if __name__ == "__main__":
    main()
