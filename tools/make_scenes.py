#!/usr/bin/env python
"""Regenerates the hw5 scene text files under scenes/ (input DATA, not code).

* practice5_1 / practice5_2 / practice5_dragon_10k are rebuilt from the compact mesh fixture
  scenes/dragon_10k_mesh.npz and checked by sha256 to be byte-identical to the course files
  the reference ships (hw5/practice5_*.txt).
* practice5_dragon_100k{,_glass,_metal,_glow}.txt: the upstream files are git-lfs blobs that
  are missing from the reference checkout (/root/reference/.MISSING_LARGE_BLOBS), so they are
  SYNTHESISED here: the 10k dragon is midpoint-subdivided (shared, slightly displaced edge
  midpoints) to 99,998 triangles inside the same Cornell box, and the material lines follow the
  file names.  bench.py says "synthetic" for these.
* scenes/rabbid.txt is the course scene `sample.txt` of the reference checkout (52 rotated boxes and
  ellipsoids lit by the background), kept verbatim as input data.
* a few small scenes of our own that exercise what the course scenes do not (emissive
  ellipsoid + rotated emissive box in the light mix, metallic, dielectric, multi-primitive
  BVH leaves).

Usage: python tools/make_scenes.py [--out scenes] [--only NAME ...]
"""
import argparse
import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SHA256 = {
    "practice5_1.txt": "ae963c11f373b70edd6c169bbe4def8df67efdbff7c4348a8b7a67a27585561e",
    "practice5_2.txt": "0af0df7d4d632283bc14d06a6f84d7a49acff42b1e806ad4dcca51003cc77821",
    "practice5_dragon_10k.txt": "7ca31e79dccb39deebc8482c946834860acd3ca7752c870e9519865051bcd014",
}

PRACTICE5_1 = """DIMENSIONS 1024 768
RAY_DEPTH 6
SAMPLES 64

BG_COLOR 1 1 1

CAMERA_POSITION 0 2 0
CAMERA_RIGHT 1 0 0
CAMERA_UP 0 1 0
CAMERA_FORWARD 0 0 -1
CAMERA_FOV_X 1.54857776

NEW_PRIMITIVE
PLANE 0 1 0
COLOR 0.25 0.25 0.5

NEW_PRIMITIVE
TRIANGLE 0 0 0 3 0 0 0 0 3
POSITION 0 0.5 -6
ROTATION 0 -0.3826834 0 0.9238795 
COLOR 1 0.125 0.125
"""

PRACTICE5_2 = """DIMENSIONS 1024 768
RAY_DEPTH 6
SAMPLES 512

BG_COLOR 0 0 0

CAMERA_POSITION 0 2 0
CAMERA_RIGHT 1 0 0
CAMERA_UP 0 1 0
CAMERA_FORWARD 0 0 -1
CAMERA_FOV_X 1.54857776

NEW_PRIMITIVE
PLANE 0 1 0
COLOR 1 1 1

NEW_PRIMITIVE
TRIANGLE 0 0 0 2 3.464 0 4 0 0
POSITION 0 3 -6
ROTATION 0.1698987 0.4530631 -0.8733046 -0.0566329 
COLOR 0 0 0
EMISSION 1 0.5 0.25

NEW_PRIMITIVE
BOX 0.5 0.5 0.5
POSITION -2 0.5 -5
ROTATION 0 0.3826834 0 0.9238795 
COLOR 0.5 0.5 1.0

NEW_PRIMITIVE
ELLIPSOID 0.7 0.7 0.7
POSITION 2 0.7 -5
COLOR 1.0 0.5 0.5

"""

CORNELL_HEADER = """DIMENSIONS {w} {h}
RAY_DEPTH {depth}
SAMPLES {spp}

BG_COLOR 0 0 0

CAMERA_POSITION 0 0 15
CAMERA_RIGHT 1 0 0
CAMERA_UP 0 1 0
CAMERA_FORWARD 0 0 -1
CAMERA_FOV_X 0.927295218

NEW_PRIMITIVE
PLANE 0 1 0
POSITION 0 -5 0
COLOR 1 1 1

NEW_PRIMITIVE
PLANE 0 0 1
POSITION 0 0 -5
COLOR 1 1 1

NEW_PRIMITIVE
PLANE 0 -1 0
POSITION 0 5 0
COLOR 1 1 1

NEW_PRIMITIVE
PLANE 1 0 0
POSITION -5 0 0
COLOR 1 0.25 0.25

NEW_PRIMITIVE
PLANE -1 0 0
POSITION 5 0 0
COLOR 0.25 1 0.25

NEW_PRIMITIVE
BOX 2 0.1 2
POSITION 0 5 0
EMISSION 2 2 2

"""

LIGHTS_MIX = """DIMENSIONS 96 64
RAY_DEPTH 5
SAMPLES 16
BG_COLOR 0.05 0.05 0.1
CAMERA_POSITION 0 1.5 6
CAMERA_RIGHT 1 0 0
CAMERA_UP 0 1 0
CAMERA_FORWARD 0 0 -1
CAMERA_FOV_X 1.2

NEW_PRIMITIVE
PLANE 0 1 0
COLOR 0.8 0.8 0.8

NEW_PRIMITIVE
PLANE 0 0 1
POSITION 0 0 -4
COLOR 0.6 0.7 0.8

NEW_PRIMITIVE
ELLIPSOID 0.4 0.25 0.3
POSITION -1.5 3 -1
ROTATION 0.1 0.2 0.3 0.9273618
COLOR 0 0 0
EMISSION 6 5 4

NEW_PRIMITIVE
BOX 0.5 0.05 0.3
POSITION 1.5 3.2 -1
ROTATION 0.2588190 0 0 0.9659258
COLOR 0 0 0
EMISSION 3 4 6

NEW_PRIMITIVE
ELLIPSOID 0.8 0.8 0.8
POSITION -1.2 0.8 -1
DIELECTRIC
IOR 1.5
COLOR 0.9 0.95 1

NEW_PRIMITIVE
BOX 0.6 0.6 0.6
POSITION 1.3 0.6 -1.5
ROTATION 0 0.3826834 0 0.9238795
METALLIC
COLOR 0.9 0.7 0.4

NEW_PRIMITIVE
TRIANGLE -1 0 0 1 0 0 0 1.5 0
POSITION 0 0 -3
COLOR 0.9 0.2 0.2

NEW_PRIMITIVE
TRIANGLE 0 0 0 1 0 0 0 1 0
POSITION -0.5 0.2 0.5
ROTATION 0 0.2588190 0 0.9659258
COLOR 0.2 0.9 0.2

NEW_PRIMITIVE
ELLIPSOID 0.3 0.5 0.3
POSITION 0.2 0.5 0.3
COLOR 0.7 0.7 0.2
"""


def fmt(x):
    return "%g" % float(x)


def tri_block(verts9, material_lines):
    return ("NEW_PRIMITIVE\nTRIANGLE " + " ".join(fmt(x) for x in verts9) + " \nPOSITION 0 0 0\n" + material_lines)


def load_mesh():
    z = np.load(os.path.join(ROOT, "scenes", "dragon_10k_mesh.npz"))
    return z["vertices"].astype(np.float32), z["triangles"].astype(np.int64)


def dragon_text(verts, tris, material_lines, w=512, h=512, spp=128, depth=6):
    out = [CORNELL_HEADER.format(w=w, h=h, spp=spp, depth=depth)]
    flat = verts[tris].reshape(-1, 9)
    for row in flat:
        out.append(tri_block(row, material_lines))
    return "".join(out)


def subdivide_midpoint(verts, tris, displace):
    """1 -> 4 midpoint subdivision of the selected mesh with shared edge midpoints.  Each new
    midpoint is pushed along the mean normal of the faces sharing the edge by
    displace * edge_length, which keeps the mesh closed while giving it the slightly curved
    facets a genuinely finer scan would have."""
    verts = verts.astype(np.float64)
    fn = np.cross(verts[tris[:, 1]] - verts[tris[:, 0]], verts[tris[:, 2]] - verts[tris[:, 0]])
    ln = np.linalg.norm(fn, axis=1, keepdims=True)
    fn = fn / np.maximum(ln, 1e-30)
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]], 0)
    ekey = np.sort(e, 1)
    uniq, inv = np.unique(ekey, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    nsum = np.zeros((len(uniq), 3))
    np.add.at(nsum, inv, np.concatenate([fn, fn, fn], 0))
    nlen = np.linalg.norm(nsum, axis=1, keepdims=True)
    nsum = nsum / np.maximum(nlen, 1e-30)
    a, b = verts[uniq[:, 0]], verts[uniq[:, 1]]
    mid = 0.5 * (a + b) + displace * np.linalg.norm(b - a, axis=1, keepdims=True) * nsum
    nv = len(verts)
    allv = np.concatenate([verts, mid], 0)
    nt = len(tris)
    m01, m12, m20 = nv + inv[:nt], nv + inv[nt:2 * nt], nv + inv[2 * nt:]
    v0, v1, v2 = tris[:, 0], tris[:, 1], tris[:, 2]
    new = np.stack([np.stack([v0, m01, m20], 1), np.stack([m01, v1, m12], 1),
                    np.stack([m20, m12, v2], 1), np.stack([m01, m12, m20], 1)], 1).reshape(-1, 3)
    return allv, new


def dragon_100k_mesh(target=100000):
    verts, tris = load_mesh()
    v1, t1 = subdivide_midpoint(verts, tris, 0.04)          # 39,968
    area = 0.5 * np.linalg.norm(np.cross(v1[t1[:, 1]] - v1[t1[:, 0]], v1[t1[:, 2]] - v1[t1[:, 0]]), axis=1)
    k = (target - len(t1)) // 3
    order = np.argsort(-area, kind="stable")
    big = np.zeros(len(t1), bool)
    big[order[:k]] = True
    # second level only on the largest facets; their sub-facets get their own (unshared
    # with unsplit neighbours) midpoints -- T-junctions are harmless to a path tracer.
    v2, t2 = subdivide_midpoint(v1, t1[big], 0.02)
    nsmall = (~big).sum()
    tris_all = np.concatenate([t1[~big], t2], 0)
    # keep file order spatially incoherent-ish like a scanned mesh: stable interleave by parent id
    parent = np.concatenate([np.nonzero(~big)[0], np.repeat(np.nonzero(big)[0], 4)])
    tris_all = tris_all[np.argsort(parent, kind="stable")]
    assert nsmall + 4 * k == len(tris_all)
    return v2.astype(np.float32), tris_all


MATERIALS = {
    "": "COLOR 0.5 0.5 1\n",
    "_glass": "DIELECTRIC\nIOR 1.5\nCOLOR 0.8 0.9 1\n",
    "_metal": "METALLIC\nCOLOR 0.9 0.75 0.4\n",
    "_glow": "COLOR 0.5 0.5 1\nEMISSION 0.2 0.3 0.6\n",
}


def write(path, text, name):
    with open(path, "w", newline="\n") as f:
        f.write(text)
    if name in SHA256:
        got = hashlib.sha256(text.encode()).hexdigest()
        if got != SHA256[name]:
            raise SystemExit("%s: sha256 %s does not match the course file %s" % (name, got, SHA256[name]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "scenes"))
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    jobs = {
        "practice5_1.txt": lambda: PRACTICE5_1,
        "practice5_2.txt": lambda: PRACTICE5_2,
        "lights_mix.txt": lambda: LIGHTS_MIX,
        "practice5_dragon_10k.txt": lambda: dragon_text(*load_mesh(), MATERIALS[""]),
    }
    cache = {}

    def mesh100k():
        if "m" not in cache:
            cache["m"] = dragon_100k_mesh()
        return cache["m"]

    for sfx, mat in MATERIALS.items():
        jobs["practice5_dragon_100k%s.txt" % sfx] = (lambda mat=mat: dragon_text(*mesh100k(), mat))
    # BASELINE.json configs[4] (metal at 3840x2160, 1024 spp) is the _metal file rendered with the
    # width/height/samples override of the C-ABI (rtc_scene_override), not a separate file.
    for name, fn in jobs.items():
        if args.only and name not in args.only:
            continue
        path = os.path.join(args.out, name)
        if os.path.exists(path) and not args.force:
            continue
        write(path, fn(), name)
        print("wrote", path)


if __name__ == "__main__":
    main()
