#!/usr/bin/env python
"""Records golden vectors from the UNMODIFIED reference (oracle/_ref/librefprobe.so, built by
oracle/Makefile from /root/reference/hw5) into tests/golden/*.npz.

Run in the development container only (the GPU box has no /root/reference; it uses the
committed fixtures).  Each fixture stores its INPUTS (rays, points, ...) next to the
reference's OUTPUTS, so a checker never needs the reference to evaluate it.

    python tools/make_golden.py            # all fixtures
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orclib  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(ROOT, "scenes")


def unit(v):
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def rays_fixture(name, stride, seed):
    """Primary rays through pixel centres + secondary rays leaving the primary hit points in
    random directions + a cloud of random rays inside the scene bounds."""
    s = orclib.Scene(orclib.ref(), os.path.join(SCENES, name + ".txt"))
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:s.height:stride, 0:s.width:stride]
    xy = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], 1).astype(np.float32)
    o, d = s.camera_rays(xy)
    pid, t, nrm, inter = s.intersect(o, d)
    hit = pid >= 0
    p = (o + t[:, None] * d)[hit]
    sd = unit(rng.normal(size=p.shape))
    so = (p + np.float32(1e-4) * sd).astype(np.float32)
    spid, st, snrm, sinter = s.intersect(so, sd)
    # random rays: origins in a box around the camera/scene, directions uniform
    lo, hi = (-4.5, -4.5, -4.5), (4.5, 4.5, 14.0)
    if name == "rabbid":
        lo, hi = (-5, -8, -8), (5, 4, 3.4)
    elif name.startswith("practice5_1") or name.startswith("practice5_2") or name == "lights_mix":
        lo, hi = (-6, 0.05, -9), (6, 5, 6)
    ro = rng.uniform(lo, hi, size=(len(so), 3)).astype(np.float32)
    rd = unit(rng.normal(size=ro.shape))
    rpid, rt, rnrm, rinter = s.intersect(ro, rd)
    # pdf of the mix distribution at the secondary origins
    pn = nrm[hit]
    pdf = s.mix_pdf(so, pn, sd)
    # ... and for directions drawn by the reference's own Distribution::Sample (half of them aim at a light)
    ld = s.ref_mix_sample(so, pn, 12345)
    lpdf = s.mix_pdf(so, pn, ld)
    tm, data = s.prims()
    aabb, links, root = s.nodes()
    np.savez_compressed(
        os.path.join(GOLD, name + "_rays.npz"),
        xy=xy, cam_o=o, cam_d=d, pid=pid, t=t, nrm=nrm, inter=inter,
        sec_o=so, sec_d=sd, sec_pid=spid, sec_t=st, sec_nrm=snrm, sec_inter=sinter,
        rnd_o=ro, rnd_d=rd, rnd_pid=rpid, rnd_t=rt, rnd_nrm=rnrm, rnd_inter=rinter,
        pdf_x=so, pdf_n=pn, pdf_d=sd, pdf=pdf, lpdf_d=ld, lpdf=lpdf,
        prim_type_material=tm, prim_data_crc=np.array([np.bitwise_xor.reduce(data.view(np.uint32).ravel())], np.uint32),
        node_links_crc=np.array([np.bitwise_xor.reduce((links.ravel().astype(np.uint64) * np.arange(1, links.size + 1, dtype=np.uint64)) & np.uint64(0xFFFFFFFF))], np.uint64),
        node_aabb_sum=aabb.astype(np.float64).sum(0), nnodes=np.array([s.nnodes]), root=np.array([root]),
        info=np.array([s.width, s.height, s.ray_depth, s.samples, s.nprims, s.nbvh, s.nnodes, s.nlights]))
    print(name, "rays:", len(o), "hit frac %.3f" % hit.mean(), "secondary", len(so))
    s.close()


def primitive_fixture():
    """Primitive::Intersect of every primitive of lights_mix / practice5_2 for random rays aimed
    roughly at the primitive (hits, misses, interior starts)."""
    out = {}
    for name in ("practice5_2", "lights_mix"):
        s = orclib.Scene(orclib.ref(), os.path.join(SCENES, name + ".txt"))
        tm, data = s.prims()
        rng = np.random.default_rng(7)
        for prim in range(s.nprims):
            pos = data[prim, 6:9]
            n = 2000
            o = (pos + rng.normal(scale=1.5, size=(n, 3))).astype(np.float32)
            o[: n // 4] = (pos + rng.normal(scale=0.2, size=(n // 4, 3))).astype(np.float32)  # many start inside
            target = pos + rng.normal(scale=0.5, size=(n, 3))
            d = unit(target - o)
            d[n // 2:] *= rng.uniform(0.2, 3.0, size=(n - n // 2, 1)).astype(np.float32)  # unnormalised directions too
            hit, t, nrm, inter = s.primitive_intersect(prim, o, d)
            key = "%s_%d" % (name, prim)
            out[key + "_o"] = o; out[key + "_d"] = d; out[key + "_hit"] = hit; out[key + "_t"] = t
            out[key + "_nrm"] = nrm; out[key + "_inter"] = inter
        s.close()
    np.savez_compressed(os.path.join(GOLD, "primitive_intersect.npz"), **out)
    print("primitive fixture:", len(out) // 6, "primitives")


def tonemap_fixture():
    s = orclib.Scene(orclib.ref(), os.path.join(SCENES, "practice5_1.txt"))
    rng = np.random.default_rng(3)
    x = np.concatenate([np.linspace(0, 4, 30000, dtype=np.float32), rng.exponential(0.5, 30000).astype(np.float32),
                        np.array([0, 1e-8, 1e-3, 0.5, 1, 10, 1e6, -0.0, -1, -0.01] * 3, np.float32)])
    x = x[: (len(x) // 3) * 3].reshape(-1, 3)
    np.savez_compressed(os.path.join(GOLD, "tonemap.npz"), rgb=x, u8=s.tonemap_u8(x))
    s.close()
    print("tonemap fixture:", x.shape)


def render_fixture(name, width, height, samples, ray_depth=-1):
    """Reference render (its own minstd streams), linear radiance before tonemapping."""
    s = orclib.Scene(orclib.ref(), os.path.join(SCENES, name + ".txt"))
    s.override(width, height, samples, ray_depth)
    img = s.ref_render_linear().reshape(height, width, 3)
    np.savez_compressed(os.path.join(GOLD, "%s_render_%dx%d_%dspp.npz" % (name, width, height, samples)),
                        mean=img, info=np.array([width, height, samples, s.ray_depth]))
    print(name, "render", img.shape, "mean", img.mean(axis=(0, 1)))
    s.close()


HW2_LIGHTS = """
NEW_LIGHT
LIGHT_INTENSITY 0.9 0.85 0.7
LIGHT_DIRECTION 0.3 1 -0.4

NEW_LIGHT
LIGHT_INTENSITY 30 20 20
LIGHT_POSITION -1 5 1
LIGHT_ATTENUATION 1 0.2 0.3

NEW_LIGHT
LIGHT_INTENSITY 4 6 12
LIGHT_POSITION 3 0.5 -4
LIGHT_ATTENUATION 1 0 0.5
"""


def hw2_lights_text(sample6):
    """The reference's hw3/sample6.txt turned into an hw2 scene (the reference ships no hw2 scene): no SAMPLES, an
    ambient term, one metallic and two dielectric primitives, a directional and two point lights."""
    out = []
    nprim = 0
    for line in sample6.split("\n"):
        w = line.split()
        if w and w[0] == "SAMPLES":
            continue
        if w and w[0] == "RAY_DEPTH":
            line = "RAY_DEPTH 5"
        out.append(line)
        if w and w[0] == "CAMERA_FOV_X":
            out.append("AMBIENT_LIGHT 0.15 0.15 0.2")
        if w and w[0] == "COLOR":
            nprim += 1
            if nprim == 3:
                out.append("METALLIC")
            elif nprim == 4:
                out += ["DIELECTRIC", "IOR 1.5"]
            elif nprim == 7:
                out += ["DIELECTRIC", "IOR 1.33"]
    return "\n".join(out) + HW2_LIGHTS


def dialect_fixture(dialect, name, width=None, height=None, samples=None):
    """Image written by the unmodified hwN program (oracle/_ref/raytracing_hwN, built by oracle/Makefile from
    /root/reference/hwN) for scenes/<name>.txt with the header lines replaced; the fixture keeps the exact scene
    text next to the 8-bit image, so a checker needs neither the scene file nor the reference.  Monte Carlo
    snapshots (hw3, hw4) also keep `u8_b`, the image of the same program at SAMPLES - 1: the programs seed their
    generator with a constant, so this is the only way to see the reference's own noise level (which is NOT ours
    when a scene has no EMISSION lines: `Color() = default` leaves Primitive::emission indeterminate in the
    reference, hw4 src/scene.cpp:39-60 then lists such primitives as lights and samples towards them)."""
    import subprocess
    import tempfile
    # "course_sampleN" = hwN's own sample scenes: read where they lie in the reference checkout (input data of the
    # reference; the fixture stores the text it was rendered from, so the repository keeps no copy of the files)
    if name == "hw2_lights":
        base = hw2_lights_text(open(os.path.join(os.environ.get("REFERENCE_ROOT", "/root/reference"), "hw3", "sample6.txt")).read())
    elif name.startswith("course_sample"):
        base = open(os.path.join(os.environ.get("REFERENCE_ROOT", "/root/reference"), "hw3", name.replace("course_", "") + ".txt")).read()
    else:
        base = open(os.path.join(SCENES, name + ".txt")).read()
    text = orclib.with_header(base, width, height, samples)

    def run(scene_text):
        with tempfile.TemporaryDirectory() as td:
            sp, op = os.path.join(td, "scene.txt"), os.path.join(td, "out.ppm")
            open(sp, "w").write(scene_text)
            subprocess.run([orclib.ref_dialect_bin(dialect), sp, op], check=True, stderr=subprocess.DEVNULL)
            return orclib.read_ppm(op).copy()

    img = run(text)
    extra = {}
    if dialect >= 3:
        mirrored = []
        for line in text.splitlines():
            w = line.split()
            if w and w[0] == "CAMERA_RIGHT":
                line = "CAMERA_RIGHT " + " ".join(repr(-float(v)) for v in w[1:4])
            mirrored.append(line)
        second = run("\n".join(mirrored) + "\n")[:, ::-1].copy()
        if dialect == 3:
            # hw3's Camera::GetToRay(float, float) adds half a pixel (hw3 src/scene.cpp:186-187): its samples cover
            # [x + 0.5, x + 1.5), so the mirrored frame flipped back is one column to the right of the plain one.
            # Column x of the plain frame pairs with column x + 1 here; the last column has no partner and
            # repeats the plain frame (its noise estimate is 0: one column of the frame).
            second = np.concatenate([second[:, 1:], img[:, -1:]], axis=1)
        extra["u8_b"] = second
    out = "hw%d_%s.npz" % (dialect, name)
    np.savez_compressed(os.path.join(GOLD, out), u8=img, text=np.frombuffer(text.encode(), np.uint8), dialect=dialect, **extra)
    print(out, img.shape, "mean", img.mean(axis=(0, 1)).round(2))


def sort_fixture():
    """std::sort / std::partition permutations on keys with many ties (what the BVH order hinges on)."""
    rng = np.random.default_rng(11)
    out = {}
    L = orclib.ref().lib
    for i, (n, kinds) in enumerate([(1, 1), (2, 1), (16, 1), (17, 1), (100, 1), (1000, 1), (9993, 1), (1000, 3), (5000, 40), (4096, 4096)]):
        key = rng.integers(0, kinds, size=n).astype(np.float32)
        perm = np.arange(n, dtype=np.int32)
        L.ref_std_sort_perm(key, perm, 0, n)
        out["sort%d_key" % i] = key; out["sort%d_perm" % i] = perm
        pred = (rng.random(n) < 0.7).astype(np.uint8)
        perm2 = np.arange(n, dtype=np.int32)
        cut = L.ref_std_partition(perm2, pred, n)
        out["part%d_pred" % i] = pred; out["part%d_perm" % i] = perm2; out["part%d_cut" % i] = np.array([cut])
    np.savez_compressed(os.path.join(GOLD, "libstdcxx_order.npz"), **out)
    print("sort fixture done")


PARSER_TEXTS = [
    # attributes given BEFORE the shape line are reset by it; the block ends at the blank line
    "DIMENSIONS 8 4\nSAMPLES 2\nRAY_DEPTH 3\nNEW_PRIMITIVE\nCOLOR 1 0 0\nIOR 9\nBOX 1 2 3\nPOSITION 1 2 3\nIOR 1.5\n\nNEW_PRIMITIVE\nPLANE 0 1 0\nPOSITION 0 0 0\nMETALLIC\n",
    # blocks back to back without blank lines; scene commands after a block lose their arguments
    "DIMENSIONS 4 4\nSAMPLES 3\nRAY_DEPTH 2\nNEW_PRIMITIVE\nELLIPSOID 1 1 1\nPOSITION 0 0 0\nNEW_PRIMITIVE\nTRIANGLE 0 0 0 1 0 0 0 1 0\nPOSITION 0 0 1\nEMISSION 1 1 1\nSAMPLES 9\nRAY_DEPTH 4\n",
    # unknown words, a stray attribute at scene level, trailing spaces, tabs, no trailing newline
    "FOO 1 2\nCOLOR 1 1 1\nDIMENSIONS 3 5  \nBG_COLOR 0.1 0.2 0.3\nCAMERA_FOV_X\t1.0\nSAMPLES 1\nRAY_DEPTH 1\nNEW_PRIMITIVE\nBOX 1 1 1\nPOSITION 0 0 0\nROTATION 0 0 0.7071068 0.7071068\nDIELECTRIC",
    # numbers in exponent / signed form, too few arguments (the rest stays as it was), extra arguments
    "DIMENSIONS 2 2\nSAMPLES 1\nRAY_DEPTH 1\nNEW_PRIMITIVE\nELLIPSOID 1e0 +2 .5\nPOSITION -1e-1 2\nCOLOR 0.5 0.25 0.125 7 7\n\nNEW_PRIMITIVE\nBOX 1 1\nPOSITION 0 0 0\n",
]


def parser_fixture():
    """What the reference's Scene::Load makes of unusual inputs (every primitive carries a POSITION:
    the reference leaves Primitive::pos uninitialised otherwise)."""
    import tempfile
    out = {}
    for i, text in enumerate(PARSER_TEXTS):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, newline="") as f:
            f.write(text)
        s = orclib.Scene(orclib.ref(), f.name)
        tm, data = s.prims()
        out["text%d" % i] = np.frombuffer(text.encode(), np.uint8)
        out["info%d" % i] = np.array([s.width, s.height, s.ray_depth, s.samples, s.nprims, s.nbvh, s.nnodes, s.nlights])
        out["tm%d" % i] = tm
        out["data%d" % i] = data
        s.close()
        os.unlink(f.name)
    np.savez_compressed(os.path.join(GOLD, "parser_quirks.npz"), **out)
    print("parser fixture:", len(PARSER_TEXTS), "texts")


# Number formats at the edge of the product reader's fast path (scene_load.cpp plain_float: plain decimal numbers are
# converted with std::from_chars, everything else goes through operator>> like the reference's whole reader).  Every
# value a failed read would leave untouched is defined by an earlier line, so nothing here is indeterminate.
NUMBER_TEXTS = [
    "DIMENSIONS 2 2\nSAMPLES 1\nRAY_DEPTH 1\n"
    "NEW_PRIMITIVE\nBOX 1. .5 +2\nPOSITION 5 6 7\nPOSITION 1.5abc 2 3\nCOLOR 0.1 0.2 0.3\nCOLOR 1E+0 2e-1 3E0\nEMISSION 0 0 0\n\n"
    "NEW_PRIMITIVE\nELLIPSOID 1 1 1\nPOSITION 1 2 3\nPOSITION 0x10 5 5\nCOLOR 0.5 0.5 0.5\nCOLOR inf 1 1\nEMISSION 1 2 3\nEMISSION nan 4 4\n\n"
    "NEW_PRIMITIVE\nBOX 1 1 1\nPOSITION 1 2 3\nPOSITION 1,5 9 9\nCOLOR 0.25 0.25 0.25\nCOLOR --1 2 2\nEMISSION 0 0 0\nIOR 1.25\nIOR 1e\n\n"
    "NEW_PRIMITIVE\nBOX 1 1 1\nPOSITION 1 2 3\nCOLOR 0.5 0.5 0.5\nCOLOR 1e400 0.75 0.75\nEMISSION 0.5 0.5 0.5\nEMISSION 1e-50 0.25 0.25\nIOR 2.5\nIOR 1e+\n\n"
    "NEW_PRIMITIVE\nBOX 1 1 1\nPOSITION -0 -.5e1 00012.5000\nCOLOR 1 1 1\nCOLOR 0.1234567890123456789 1e-3 123456789\nEMISSION 0 0 0\nROTATION 0 0 0 1\nROTATION 0.5 .5 5e-1 +0.5\n",
    # CRLF line ends: the carriage return is white space to operator>>, a line of just \r ends a block
    "DIMENSIONS 2 2\r\nSAMPLES 1\r\nRAY_DEPTH 1\r\nNEW_PRIMITIVE\r\nBOX 1 2 3\r\nPOSITION 0 0 0\r\nCOLOR 0.5 0.25 1\r\nEMISSION 0 0 0\r\nMETALLIC\r\n\r\n"
    "NEW_PRIMITIVE\r\nPLANE 0 1 0\r\nPOSITION 0 -1 0\r\nCOLOR 1 1 1\r\nEMISSION 0 0 0\r\n",
]


def parser_number_fixture():
    import tempfile
    out = {}
    for i, text in enumerate(NUMBER_TEXTS):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, newline="") as f:
            f.write(text)
        s = orclib.Scene(orclib.ref(), f.name)
        tm, data = s.prims()
        out["text%d" % i] = np.frombuffer(text.encode(), np.uint8)
        out["info%d" % i] = np.array([s.width, s.height, s.ray_depth, s.samples, s.nprims, s.nbvh, s.nnodes, s.nlights])
        out["tm%d" % i] = tm
        out["data%d" % i] = data
        s.close()
        os.unlink(f.name)
    np.savez_compressed(os.path.join(GOLD, "parser_numbers.npz"), **out)
    print("parser number fixture:", len(NUMBER_TEXTS), "texts")


def headline():
    """The headline scenes pinned to the reference itself (round 2): golden rays of the three 100k dragons (primary,
    first-bounce secondary and random rays through librefprobe -> Scene::RayIntersection) and one larger converged
    render of practice5_dragon_100k.  ~25 minutes on 8 cores: the reference visits ~10^4 nodes per ray here."""
    rays_fixture("practice5_dragon_100k", 2, 6)          # 65,536 primary rays + as many secondary and random ones
    rays_fixture("practice5_dragon_100k_glass", 4, 7)
    rays_fixture("practice5_dragon_100k_metal", 4, 8)
    render_fixture("practice5_dragon_100k", 96, 96, 512)


def main():
    if not orclib.have_ref():
        raise SystemExit("oracle/_ref/librefprobe.so missing: run `make -C oracle` where /root/reference exists")
    os.makedirs(GOLD, exist_ok=True)
    sort_fixture()
    parser_fixture()
    tonemap_fixture()
    primitive_fixture()
    rays_fixture("practice5_1", 8, 1)
    rays_fixture("practice5_2", 8, 2)
    rays_fixture("lights_mix", 1, 3)
    rays_fixture("practice5_dragon_10k", 4, 4)
    rays_fixture("rabbid", 2, 5)
    render_fixture("practice5_1", 64, 48, 256)
    render_fixture("practice5_2", 64, 48, 1024)
    render_fixture("lights_mix", 48, 32, 1024)
    render_fixture("practice5_dragon_10k", 64, 64, 256)
    render_fixture("practice5_dragon_10k", 128, 128, 512)
    render_fixture("rabbid", 88, 88, 256)
    for sfx in ("", "_glass", "_metal"):
        render_fixture("practice5_dragon_100k" + sfx, 48, 48, 128)
    dialects()


def dialects():
    """The four earlier snapshots only write images: deterministic ones at the scene's own size, Monte Carlo
    ones converged at a reduced size."""
    dialect_fixture(1, "course_sample6")
    dialect_fixture(2, "hw2_lights")
    dialect_fixture(2, "hw2_glass")        # nested dielectrics, total internal reflection, metallic wall, RAY_DEPTH 8
    dialect_fixture(1, "course_sample3")
    dialect_fixture(1, "course_sample5")
    dialect_fixture(3, "course_sample6", 103, 133, 2048)
    dialect_fixture(4, "course_sample6", 103, 133, 2048)
    dialect_fixture(3, "course_sample4", 64, 64, 4096)
    dialect_fixture(4, "course_sample4", 64, 64, 2048)
    dialect_fixture(3, "course_sample3", 64, 64, 2048)   # metallic box + emissive box
    dialect_fixture(4, "course_sample3", 64, 64, 1024)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dialects":
        dialects()
    elif len(sys.argv) > 1 and sys.argv[1] == "headline":
        headline()
    elif len(sys.argv) > 1 and sys.argv[1] == "parser_numbers":
        parser_number_fixture()
    else:
        main()
