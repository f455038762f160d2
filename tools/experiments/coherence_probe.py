"""How much would ray sorting help k_traverse?  Traces the same 4 M secondary rays in three orders
(pixel order / shuffled / sorted by origin cell + direction octant) through rtc_intersect; run under
`ncu --metrics gpu__time_duration.sum -k regex:k_traverse` to read the three kernel times."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import raytracing_course_b200 as rtc

s = rtc.Scene(path=os.path.join(ROOT, "scenes", "practice5_dragon_100k.txt"), device=0)
rng = np.random.default_rng(0)
reps = 16
ys, xs = np.mgrid[0:s.height, 0:s.width]
xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)
xy = np.repeat(xy, reps, axis=0) + rng.random((len(xy) * reps, 2), dtype=np.float32)
o, d = s.cam.GetToRay(xy)
pid, t, n, inter = s.RayIntersection(o, d)
hit = pid >= 0
p = (o + t[:, None] * d)[hit]
nn = n[hit]
g = rng.normal(size=p.shape).astype(np.float32)
g /= np.linalg.norm(g, axis=1, keepdims=True)
dirs = g + nn                                   # cosine lobe like SampleCosine
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
org = (p + 1e-4 * dirs).astype(np.float32)
print("rays", len(org))


def key(org, dirs, bits):
    lo, hi = org.min(0), org.max(0)
    q = np.clip(((org - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.uint32), 0, (1 << bits) - 1)
    k = np.zeros(len(org), np.uint64)
    for b in range(bits):
        for a in range(3):
            k |= ((q[:, a] >> b) & 1).astype(np.uint64) << np.uint64(3 * b + a)
    octant = (dirs[:, 0] > 0).astype(np.uint64) | ((dirs[:, 1] > 0).astype(np.uint64) << np.uint64(1)) | ((dirs[:, 2] > 0).astype(np.uint64) << np.uint64(2))
    return (k << np.uint64(3)) | octant


orders = {"pixel": np.arange(len(org)), "shuffled": rng.permutation(len(org)),
          "sorted5": np.argsort(key(org, dirs, 5), kind="stable"), "sorted3": np.argsort(key(org, dirs, 3), kind="stable"),
          "octant_only": np.argsort(key(org, dirs, 0), kind="stable")}
ref = None
for name, perm in orders.items():
    r = s.RayIntersection(org[perm], dirs[perm])
    ids = np.empty_like(r[0]); ids[perm] = r[0]
    if ref is None: ref = ids
    print(name, "same ids:", bool(np.array_equal(ids, ref)))
