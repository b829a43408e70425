"""Generates tests/golden/*.json|npy from the reference checkout (run in the build container,
where /root/reference is mounted; the GPU box only sees the committed outputs).

Sources (paths relative to /root/reference):
  radius_grid5x5       test/topology.jl:46-52   5x5 grid, spacing 0.1, r = 0.15 (diagonal 0.1414 < r: 8-neighbourhood)
  circle_k3            test/neighbors.jl:36-56  20 points on the unit circle, k = 3 including self
  compute_force        test/repel.jl:117-170    closed forms evaluated at the test's sample points
  closest_point        test/octree_geometric.jl:6-49
  stl face centres     test/data/cavity.stl, bifurcation.stl (binary STL, float32) -> real-geometry inputs
"""
import json, os, struct, sys
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def stl_face_centres(path, limit=None):
    with open(path, "rb") as f:
        f.read(80)
        n = struct.unpack("<I", f.read(4))[0]
        rec = np.frombuffer(f.read(50 * n), dtype=np.uint8).reshape(n, 50)
    tri = rec[:, 12:48].copy().view(np.float32).reshape(n, 3, 3)
    c = ((tri[:, 0] + tri[:, 1]) + tri[:, 2]) / np.float32(3)
    c = np.unique(c, axis=0)  # exact duplicates would take _safe_direction's random branch
    return c[:limit] if limit else c


g = {}
grid = [[i * 0.1, j * 0.1] for i in range(5) for j in range(5)]
rows = []
for a, (xa, ya) in enumerate(grid):
    rows.append([b + 1 for b, (xb, yb) in enumerate(grid) if b != a and abs(round((xa - xb) / 0.1)) <= 1 and abs(round((ya - yb) / 0.1)) <= 1])
g["radius_grid5x5"] = {"points": grid, "radius": 0.15, "rows": rows, "source": "test/topology.jl:46-52"}
N = 20
theta = np.linspace(0, 2 * np.pi, N + 1)[:-1]
g["circle_k3"] = {"points": np.stack([np.cos(theta), np.sin(theta)], 1).tolist(),
                  "sets": [sorted([i + 1, (i - 1) % N + 1, (i + 1) % N + 1]) for i in range(N)], "source": "test/neighbors.jl:36-56"}
fu = [0.0, 0.3, 0.5, 0.7, 0.79, 0.8, 0.99, 1.0, 1.5, 2.0, 10.0]
g["compute_force"] = {
    "u": fu, "beta": 0.2,
    "inverse": [1 / (u * u + 0.2) ** 2 for u in fu],
    "equilibrium": [(1 - u * u) / (u * u + 0.2) ** 2 for u in fu],
    "clipped_u0_1": [max((1 - u * u) / (u * u + 0.2) ** 2, 0.0) for u in fu],
    "clipped_u0_0.8": [max((0.8 * 0.8 - u * u) / (u * u + 0.2) ** 2, 0.0) for u in fu],
    "strong_gamma3": [(1 - u * u) / (u * u + 0.2) ** 3 for u in fu],
    "source": "test/repel.jl:117-170"}
g["closest_point"] = {"tri": [[0, 0, 0], [1, 0, 0], [0, 1, 0]],
                      "cases": [([0.25, 0.25, 1.0], [0.25, 0.25, 0.0]), ([-1, -1, 0], [0, 0, 0]), ([2, -1, 0], [1, 0, 0]), ([-1, 2, 0], [0, 1, 0]),
                                ([0.5, -0.5, 0], [0.5, 0, 0]), ([-0.5, 0.5, 0], [0, 0.5, 0]), ([0.3, 0.3, 0], [0.3, 0.3, 0])],
                      "source": "test/octree_geometric.jl:6-49"}
with open(os.path.join(OUT, "reference_known_answers.json"), "w") as f:
    json.dump(g, f, indent=1)
if os.path.isdir(REF):
    np.save(os.path.join(OUT, "cavity_face_centres_f32.npy"), stl_face_centres(os.path.join(REF, "test/data/cavity.stl")))
    np.save(os.path.join(OUT, "bifurcation_face_centres_f32.npy"), stl_face_centres(os.path.join(REF, "test/data/bifurcation.stl"), 8000))
print({k: os.path.getsize(os.path.join(OUT, k)) for k in os.listdir(OUT)})
