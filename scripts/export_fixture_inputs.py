"""Writes the inputs of the reference fixtures: the clouds the parity tests use, as raw
little-endian arrays + manifest.json under tests/golden/reference/inputs/. scripts/make_reference_fixtures.jl (run with
Julia and the unmodified reference) turns them into neighbour tables and 10-iteration positions;
tests/test_reference_fixtures.py compares the oracle and the device against those when they are present."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference", "inputs")


def cases():
    rng = np.random.default_rng(2024)
    out = []
    for dt in (np.float32, np.float64):
        tag = np.dtype(dt).name
        out.append(dict(name=f"knn_u3_100k_{tag}", kind="knn", k=21, pts=synth.uniform_cube(100_000, dt)))
        out.append(dict(name=f"knn_u2_50k_{tag}", kind="knn", k=21, pts=rng.random((50_000, 2)).astype(dt)))
        out.append(dict(name=f"radius_u2_50k_{tag}", kind="radius", r=float(2.5 * 50_000 ** -0.5), pts=rng.random((50_000, 2)).astype(dt)))
        snap = rng.random((6000, 3)).astype(dt)
        h = 6000 ** (-1 / 3)
        out.append(dict(name=f"repel_const_6k_{tag}", kind="repel", pts=snap, n_fixed=800, spacing=dict(kind="constant", a=h), beta=0.2, k=21,
                        max_iters=10, alpha_max=h / 20, alpha_lo=h / 2000))
        gp, nw, hw = synth.graded_cube(60_000, dt)
        out.append(dict(name=f"repel_graded_60k_{tag}", kind="repel", pts=gp, n_fixed=int(nw),
                        spacing=dict(kind="boundary_layer", a=hw, b=4 * hw, c=0.2), beta=0.2, k=21, max_iters=10, alpha_max=hw / 20, alpha_lo=hw / 2000))
    grid = np.array([[i * 0.1, j * 0.1] for i in range(5) for j in range(5)])
    out.append(dict(name="radius_grid5x5_float64", kind="radius", r=0.15, pts=grid))                    # test/topology.jl:46-52
    for stl in ("cavity", "bifurcation"):
        p = np.load(os.path.join(ROOT, "tests", "golden", f"{stl}_face_centres_f32.npy"))
        out.append(dict(name=f"knn_{stl}_float32", kind="knn", k=21, pts=p))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    for c in cases():
        pts = np.ascontiguousarray(c.pop("pts"))
        fn = c["name"] + ".bin"
        pts.astype(pts.dtype.newbyteorder("<")).tofile(os.path.join(OUT, fn))
        manifest.append(dict(c, file=fn, dtype=pts.dtype.name, n=int(pts.shape[0]), d=int(pts.shape[1])))
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(f"{len(manifest)} cases -> {OUT}")


if __name__ == "__main__":
    main()
