"""Does longest-first ordering shorten the tail of the persistent k_traverse?  Traces 1 M secondary
rays in (a) pixel order, (b) descending chord length through the BVH bounds, (c) ascending; run under
`ncu --metrics gpu__time_duration.sum -k regex:k_traverse`."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import raytracing_course_b200 as rtc

s = rtc.Scene(path=os.path.join(ROOT, "scenes", "practice5_dragon_100k.txt"), device=0)
rng = np.random.default_rng(0)
ys, xs = np.mgrid[0:s.height, 0:s.width]
xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)
xy = np.repeat(xy, 4, axis=0) + rng.random((len(xy) * 4, 2), dtype=np.float32)
o, d = s.cam.GetToRay(xy)
pid, t, n, inter = s.RayIntersection(o, d)
hit = pid >= 0
p = (o + t[:, None] * d)[hit]; nn = n[hit]
g = rng.normal(size=p.shape).astype(np.float32); g /= np.linalg.norm(g, axis=1, keepdims=True)
dirs = g + nn; dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
org = (p + 1e-4 * dirs).astype(np.float32)
aabb, links, root = s.nodes()
lo, hi = aabb[root, :3], aabb[root, 3:]
inv = 1.0 / dirs
t1 = (lo - org) * inv; t2 = (hi - org) * inv
tn = np.minimum(t1, t2).max(1); tf = np.maximum(t1, t2).min(1)
chord = np.where((tn <= tf) & (tf >= 0), tf - np.maximum(tn, 0), 0).astype(np.float32)
print("rays", len(org), "chord>0", (chord > 0).mean())
orders = {"pixel": np.arange(len(org)), "long_first": np.argsort(-chord, kind="stable"), "short_first": np.argsort(chord, kind="stable"),
          "two_bins": np.argsort(chord < np.median(chord[chord > 0]), kind="stable")}
for name, perm in orders.items():
    for rep in range(2):
        r = s.RayIntersection(org[perm], dirs[perm])
    print(name, "done")
