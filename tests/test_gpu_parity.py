"""GPU (-m gpu): the CUDA path, called through the C-ABI, against the oracle and against the golden
vectors recorded from the unmodified reference.

Tolerances (north_star): primitive ids identical on >= 99.9 % of rays, hit distance within 1e-4
relative on the agreeing rays; u8 output within 1 LSB; rendered radiance either sample-exact
against the oracle's Philox streams (same streams, float rounding differences only) or
statistically equal to the reference's own render."""
import os
import subprocess

import numpy as np
import pytest

import orclib
from conftest import golden, scene_path

pytestmark = pytest.mark.gpu

RAY_SCENES = ["practice5_1", "practice5_2", "lights_mix", "practice5_dragon_10k", "rabbid"]
HEADLINE_SCENES = ["practice5_dragon_100k", "practice5_dragon_100k_glass", "practice5_dragon_100k_metal"]
ID_AGREE = 0.999
T_REL = 1e-4


def check_hits(got, want, id_agree=ID_AGREE, strict_t=False):
    pid, t, nrm, inter = got
    wpid, wt, wnrm, winter = want
    same = pid == wpid
    assert same.mean() >= id_agree, "ids agree on %.5f %% only (%d differ)" % (100 * same.mean(), (~same).sum())
    hit = same & (wpid >= 0)
    rel = np.abs(t[hit] - wt[hit]) / np.maximum(np.abs(wt[hit]), 1e-20) if hit.any() else np.zeros(1)
    if strict_t:
        assert rel.max() <= T_REL, rel.max()
    else:
        # scattered rays include directions almost parallel to a (triangle) plane, where
        # t = -dot(o,n)/dot(d,n) cancels catastrophically and FMA contraction moves it: 99.9 % within
        # 1e-4, every one within 5 %
        assert np.quantile(rel, 0.999) <= T_REL and rel.max() <= 5e-2, (np.quantile(rel, 0.999), rel.max())
    nerr = np.abs(nrm[hit] - wnrm[hit]).max(axis=1) if hit.any() else np.zeros(1)
    # normals of grazing ellipsoid hits inherit the ill-conditioned root: 99.9 % within 1e-4, all within 1e-2
    assert np.quantile(nerr, 0.999) <= 1e-4 and nerr.max() <= 1e-2, (np.quantile(nerr, 0.999), nerr.max())
    assert (inter[hit] == winter[hit]).mean() >= 0.9999
    miss = same & (wpid < 0)
    assert np.all(t[miss] == 0)
    return same.mean()


@pytest.mark.parametrize("name", RAY_SCENES)
@pytest.mark.parametrize("mode", [0, 1])
def test_ray_intersection_vs_reference_golden(rtc, gpu_scenes, name, mode):
    g = golden(name + "_rays")
    s = gpu_scenes(name)
    for kind, pre in (("cam", ""), ("sec", "sec_"), ("rnd", "rnd_")):
        got = s.RayIntersection(g[kind + "_o"], g[kind + "_d"], mode)
        check_hits(got, (g[pre + "pid"], g[pre + "t"], g[pre + "nrm"], g[pre + "inter"]), strict_t=(kind == "cam"))


@pytest.mark.parametrize("name", HEADLINE_SCENES)
def test_headline_scenes_vs_reference_golden(rtc, gpu_scenes, name):
    """The 100k dragons against rays recorded from the compiled reference itself (tools/make_golden.py headline):
    65,536 (16,384) primary rays, as many first-bounce secondary rays leaving the primary hit points and random rays.
    The index traversal must give the reference's primitive on EVERY one of them (zero disagreements: this bounds the
    documented hole of DESIGN.md, a ray grazing an ancestor box within an ulp, on the scene's real rays), t within
    1e-4 relative, and so must the node-by-node twin."""
    g = golden(name + "_rays")
    s = gpu_scenes(name)
    o, d = s.cam.GetToRay(g["xy"])
    assert np.array_equal(o.view(np.uint32), g["cam_o"].view(np.uint32))
    assert np.array_equal(d.view(np.uint32), g["cam_d"].view(np.uint32))
    for mode in (0, 1):
        for kind, pre in (("cam", ""), ("sec", "sec_"), ("rnd", "rnd_")):
            pid, t, nrm, inter = s.RayIntersection(g[kind + "_o"], g[kind + "_d"], mode)
            want = g[pre + "pid"]
            assert np.array_equal(pid, want), (name, mode, kind, int((pid != want).sum()))
            hit = want >= 0
            rel = np.abs(t[hit] - g[pre + "t"][hit]) / np.maximum(np.abs(g[pre + "t"][hit]), 1e-12)
            assert rel.max() <= T_REL, (name, mode, kind, rel.max())
            assert np.array_equal(inter[hit], g[pre + "inter"][hit])


@pytest.mark.parametrize("name", ["practice5_dragon_10k", "practice5_dragon_100k", "practice5_dragon_100k_glass"])
def test_primary_hits_full_frame_vs_oracle(rtc, gpu_scenes, oracle_scenes, name):
    """Every pixel centre of the frame (BASELINE parity criterion): ids >= 99.9 %, t within 1e-4."""
    s, a = gpu_scenes(name), oracle_scenes(name)
    ys, xs = np.mgrid[0:s.height, 0:s.width]
    xy = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], 1).astype(np.float32)
    o, d = s.cam.GetToRay(xy)
    ao, ad = a.camera_rays(xy)
    assert np.array_equal(o.view(np.uint32), ao.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), ad.view(np.uint32))
    want = a.intersect(o, d)
    agree = check_hits(s.RayIntersection(o, d, rtc.TRAVERSAL_INDEX), want, strict_t=True)
    assert agree >= 0.9999
    sub = slice(0, None, 16)  # the node-by-node twin is slow: every 16th ray
    check_hits(s.RayIntersection(o[sub], d[sub], rtc.TRAVERSAL_REFTREE), tuple(w[sub] for w in want))


def test_index_and_reference_tree_traversals_agree_on_scattered_rays(rtc, gpu_scenes, oracle_scenes):
    """Secondary-like rays (random origins inside the box, random directions) at 100k triangles."""
    name = "practice5_dragon_100k"
    s, a = gpu_scenes(name), oracle_scenes(name)
    rng = np.random.default_rng(21)
    n = 200000
    o = rng.uniform((-4.9, -4.9, -4.9), (4.9, 4.9, 6.0), size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    fast = s.RayIntersection(o, d, rtc.TRAVERSAL_INDEX)
    slow = s.RayIntersection(o, d, rtc.TRAVERSAL_REFTREE)
    check_hits(fast, slow)
    k = 20000
    check_hits(tuple(x[:k] for x in fast), a.intersect(o[:k], d[:k]))


def test_empty_and_tiny_batches(rtc, gpu_scenes):
    s = gpu_scenes("practice5_2")
    pid, t, nrm, inter = s.RayIntersection(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert pid.shape == (0,)
    pid, t, nrm, inter = s.RayIntersection(np.array([[0, 2, 0]], np.float32), np.array([[0, 1, 0]], np.float32))
    assert pid[0] == -1 and t[0] == 0  # straight up: misses everything


def test_primitive_intersect_vs_reference_golden(rtc, gpu_scenes):
    g = golden("primitive_intersect")
    for name in ("practice5_2", "lights_mix"):
        s = gpu_scenes(name)
        for prim in range(s.nprims):
            key = "%s_%d" % (name, prim)
            hit, t, nrm, inter = s.PrimitiveIntersect(prim, g[key + "_o"], g[key + "_d"])
            same = hit == g[key + "_hit"]
            assert same.mean() >= 0.999, key
            m = same & (hit == 1)
            rel = np.abs(t[m] - g[key + "_t"][m]) / np.maximum(np.abs(g[key + "_t"][m]), 1e-20)
            # rays that graze an ellipsoid have an ill-conditioned root: allow 1e-3 there
            assert np.quantile(rel, 0.999) <= T_REL and rel.max() <= 1e-2, (key, rel.max())
            assert (inter[m] == g[key + "_inter"][m]).mean() >= 0.999


@pytest.mark.parametrize("name", RAY_SCENES)
def test_camera_rays_bit_exact(rtc, gpu_scenes, name):
    g = golden(name + "_rays")
    o, d = gpu_scenes(name).cam.GetToRay(g["xy"])
    assert np.array_equal(o.view(np.uint32), g["cam_o"].view(np.uint32))
    assert np.array_equal(d.view(np.uint32), g["cam_d"].view(np.uint32))


@pytest.mark.parametrize("name", RAY_SCENES)
def test_mix_pdf_vs_reference_golden(rtc, gpu_scenes, name):
    g = golden(name + "_rays")
    pdf = gpu_scenes(name).mix_distrib.Pdf(g["pdf_x"], g["pdf_n"], g["pdf_d"])
    rel = np.abs(pdf - g["pdf"]) / np.maximum(np.abs(g["pdf"]), 1e-12)
    assert np.quantile(rel, 0.999) <= 1e-4, np.quantile(rel, 0.999)


@pytest.mark.parametrize("name", RAY_SCENES)
def test_mix_pdf_towards_lights_vs_reference_golden(rtc, gpu_scenes, name):
    """As above for directions drawn by the reference's own Distribution::Sample (half of them hit a
    light).  Grazing hits have |cos| ~ 1e-3 in the denominator: 99 % within 1e-4, 99.9 % within 1e-2."""
    g = golden(name + "_rays")
    pdf = gpu_scenes(name).mix_distrib.Pdf(g["pdf_x"], g["pdf_n"], g["lpdf_d"])
    rel = np.abs(pdf - g["lpdf"]) / np.maximum(np.abs(g["lpdf"]), 1e-12)
    assert np.quantile(rel, 0.99) <= 1e-4 and np.quantile(rel, 0.999) <= 1e-2, (np.quantile(rel, 0.99), np.quantile(rel, 0.999))


@pytest.mark.parametrize("name", ["lights_mix", "practice5_dragon_10k"])
def test_mix_sample_vs_oracle_same_streams(rtc, gpu_scenes, oracle_scenes, name):
    g = golden(name + "_rays")
    x, n = g["pdf_x"], g["pdf_n"]
    got = gpu_scenes(name).mix_distrib.Sample(x, n, seed=5, sample=3, bounce=2)
    want = oracle_scenes(name).mix_sample(x, n, 5, 3, 2)
    close = np.abs(got - want).max(axis=1) <= 1e-4
    assert close.mean() >= 0.999, close.mean()
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)


def test_tonemap_vs_reference_golden(rtc, gpu_scenes):
    g = golden("tonemap")
    out = gpu_scenes("practice5_1").ToUInts(g["rgb"])
    diff = np.abs(out.astype(np.int32) - g["u8"].astype(np.int32))
    assert diff.max() <= 1
    assert (diff == 0).mean() >= 0.995


def squash(x):
    return x / (1.0 + x)


@pytest.mark.parametrize("name,w,h,spp", [("practice5_1", 64, 48, 4), ("practice5_2", 64, 48, 8), ("lights_mix", 96, 64, 8),
                                          ("practice5_dragon_10k", 64, 64, 4), ("practice5_dragon_100k_glass", 40, 40, 2),
                                          ("practice5_dragon_100k_metal", 40, 40, 2), ("practice5_dragon_100k_glow", 40, 40, 2), ("rabbid", 88, 88, 4)])
def test_render_sample_exact_vs_oracle(rtc, oracle_lib, name, w, h, spp):
    """Same Philox streams on both sides: per-pixel radiance sums agree to float rounding except
    where a path crosses a discontinuity differently (a handful of pixels)."""
    s = rtc.Scene(path=scene_path(name), device=0)
    s.override(w, h, spp)
    got = s.RenderSum(seed=9, sample_begin=2, sample_count=spp).reshape(-1, 3)
    cnt = s.counters()
    a = orclib.Scene(oracle_lib, scene_path(name))
    a.override(w, h, spp)
    want, paths, rays = a.render_sum(9, 2, spp)
    a.close()
    assert cnt["paths"] == paths == w * h * spp
    assert abs(cnt["rays"] - rays) <= 1e-3 * rays
    ok = np.isfinite(want).all(1) & np.isfinite(got).all(1)
    rel = np.abs(got - want).max(1) / (np.abs(want).max(1) + 1e-3)
    good = (rel <= 2e-3) & ok
    assert good.mean() >= 0.99, good.mean()
    assert abs(np.median(squash(got[ok])) - np.median(squash(want[ok]))) < 1e-3
    # splitting the sample range (what spp-sharding over GPUs does) reproduces the same sums
    half = s.RenderSum(seed=9, sample_begin=2, sample_count=spp // 2).reshape(-1, 3) + \
        s.RenderSum(seed=9, sample_begin=2 + spp // 2, sample_count=spp - spp // 2).reshape(-1, 3)
    ok2 = np.isfinite(half).all(1) & ok
    assert np.allclose(half[ok2], got[ok2], rtol=1e-4, atol=1e-5)
    s.close()


@pytest.mark.parametrize("name,w,h,spp", [("practice5_1", 64, 48, 256), ("practice5_2", 64, 48, 1024), ("lights_mix", 48, 32, 1024),
                                          ("practice5_dragon_10k", 64, 64, 256), ("practice5_dragon_10k", 128, 128, 512), ("rabbid", 88, 88, 256),
                                          ("practice5_dragon_100k", 48, 48, 128), ("practice5_dragon_100k", 96, 96, 512),
                                          ("practice5_dragon_100k_glass", 48, 48, 128),
                                          ("practice5_dragon_100k_metal", 48, 48, 128)])
@pytest.mark.parametrize("mode", [0, 1])
def test_render_statistically_matches_reference(rtc, name, w, h, spp, mode):
    """Converged-image criterion: against the REFERENCE's own render (its minstd streams).  The
    RMSE to the reference must not exceed the RMSE between two independent renders of ours by more
    than 20 %, and the mean difference must be within 4 standard errors (no bias)."""
    g = golden("%s_render_%dx%d_%dspp" % (name, w, h, spp))
    ref = g["mean"].astype(np.float64)
    s = rtc.Scene(path=scene_path(name), device=0)
    s.override(w, h, spp)
    s.set_traversal(mode)
    a = s.RenderSum(seed=101, sample_count=spp).astype(np.float64) / spp
    b = s.RenderSum(seed=202, sample_count=spp).astype(np.float64) / spp
    s.close()
    ok = np.isfinite(ref) & np.isfinite(a) & np.isfinite(b)
    assert ok.mean() > 0.999
    sa, sb, sr = squash(a[ok]), squash(b[ok]), squash(ref[ok])
    noise = np.sqrt(np.mean((sa - sb) ** 2))
    err = np.sqrt(np.mean((sa - sr) ** 2))
    assert err <= 1.2 * noise + 1e-4, (err, noise)
    diff = sa - sr
    se = diff.std() / np.sqrt(diff.size) + 1e-12
    assert abs(diff.mean()) < 4 * se + 2e-4, (diff.mean(), se)


def test_render_u8_and_ppm_roundtrip(rtc, tmp_path):
    s = rtc.Scene(path=scene_path("practice5_2"), device=0)
    s.override(96, 72, 16)
    img = s.Render(seed=1)
    assert img.shape == (72, 96, 3) and img.dtype == np.uint8
    out = tmp_path / "out.ppm"
    s.RenderPPM(str(out), seed=1)
    raw = out.read_bytes()
    header = b"P6\n96 72\n255\n"
    assert raw.startswith(header) and len(raw) == len(header) + 96 * 72 * 3
    body = np.frombuffer(raw[len(header):], np.uint8).reshape(72, 96, 3)
    assert (np.abs(body.astype(int) - img.astype(int)) <= 1).mean() > 0.999  # atomics reorder float sums
    # against u8 of the linear sums through the reference's tonemap restatement
    lin = s.RenderSum(seed=1, sample_count=16) / 16.0
    s.close()


def test_cli_run_sh(rtc, tmp_path):
    """run.sh <scene> <out.ppm> end to end on the GPU."""
    from conftest import ROOT
    out = tmp_path / "cli.ppm"
    env = dict(os.environ, RTC_SAMPLES="4", RTC_WIDTH="128", RTC_HEIGHT="96")
    r = subprocess.run([os.path.join(ROOT, "run.sh"), scene_path("practice5_1"), str(out)], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    assert raw.startswith(b"P6\n128 96\n255\n") and len(raw) == len(b"P6\n128 96\n255\n") + 128 * 96 * 3


def test_batching_does_not_change_the_image(rtc):
    s = rtc.Scene(path=scene_path("lights_mix"), device=0)
    s.override(96, 64, 8)
    a = s.RenderSum(seed=3, sample_count=8)
    s.set_batch_paths(5000)  # many small wavefront batches, cutting through sample planes
    b = s.RenderSum(seed=3, sample_count=8)
    s.close()
    ok = np.isfinite(a) & np.isfinite(b)
    assert np.allclose(a[ok], b[ok], rtol=1e-4, atol=1e-5)


def test_full_size_properties_dragon_100k(rtc, gpu_scenes):
    """BASELINE-size frame (512x512, depth 6): size-independent properties -- every path is counted,
    rays per path within [1, depth], energy bounded by the light, image deterministic per seed."""
    s = gpu_scenes("practice5_dragon_100k")
    s.override(samples=4)
    s.reset_counters()
    a = s.RenderSum(seed=7, sample_count=4)
    c = s.counters()
    assert c["paths"] == 512 * 512 * 4
    assert c["paths"] <= c["rays"] <= 6 * c["paths"]
    assert c["fallback_rays"] <= 1e-4 * c["rays"]
    b = s.RenderSum(seed=7, sample_count=4)
    ok = np.isfinite(a) & np.isfinite(b)
    assert ok.mean() > 0.9999
    assert np.allclose(a[ok], b[ok], rtol=1e-4, atol=1e-5)
    assert (a[ok] >= 0).all()
    s.override(samples=128)


# ---------------------------------------------------------------------------------------------
# edge cases: scenes built in the test, compared with the oracle on the same text
EDGE_SCENES = {
    "planes_only": "DIMENSIONS 33 17\nSAMPLES 3\nRAY_DEPTH 4\nBG_COLOR 0.2 0.3 0.4\nCAMERA_POSITION 0 1 3\nCAMERA_RIGHT 1 0 0\nCAMERA_UP 0 1 0\n"
                   "CAMERA_FORWARD 0 0 -1\nCAMERA_FOV_X 1.2\nNEW_PRIMITIVE\nPLANE 0 1 0\nCOLOR 0.8 0.8 0.8\nNEW_PRIMITIVE\nPLANE 0 0 1\n"
                   "POSITION 0 0 -3\nROTATION 0 0.1736482 0 0.9848078\nCOLOR 0.9 0.2 0.2\nEMISSION 0.5 0.5 0.5\n",
    "one_sample_depth_one": "DIMENSIONS 31 9\nSAMPLES 1\nRAY_DEPTH 1\nBG_COLOR 1 1 1\nCAMERA_POSITION 0 0 4\nCAMERA_RIGHT 1 0 0\nCAMERA_UP 0 1 0\n"
                            "CAMERA_FORWARD 0 0 -1\nCAMERA_FOV_X 1.0\nNEW_PRIMITIVE\nELLIPSOID 1 0.5 0.7\nCOLOR 0.5 0.6 0.7\nEMISSION 0.1 0.2 0.3\n",
    # five identical overlapping ellipsoids: the SAH refuses to split them -> one reference leaf with 5 primitives
    "multi_primitive_leaf": "DIMENSIONS 40 30\nSAMPLES 4\nRAY_DEPTH 5\nBG_COLOR 0.1 0.1 0.1\nCAMERA_POSITION 0 0 5\nCAMERA_RIGHT 1 0 0\nCAMERA_UP 0 1 0\n"
                            "CAMERA_FORWARD 0 0 -1\nCAMERA_FOV_X 1.0\n" +
                            "".join("NEW_PRIMITIVE\nELLIPSOID 1 1 1\nPOSITION 0 0 0\nCOLOR 0.%d 0.5 0.5\n\n" % (i + 2) for i in range(5)) +
                            "NEW_PRIMITIVE\nBOX 0.3 0.3 0.3\nPOSITION 2 0 0\nEMISSION 5 5 5\n\nNEW_PRIMITIVE\nPLANE 0 1 0\nPOSITION 0 -1 0\nCOLOR 0.7 0.7 0.7\n",
    # glass ball in front of a metal box and a rotated emissive ellipsoid: deep specular chains
    "specular_chain": "DIMENSIONS 48 32\nSAMPLES 6\nRAY_DEPTH 12\nBG_COLOR 0.3 0.4 0.5\nCAMERA_POSITION 0 1 5\nCAMERA_RIGHT 1 0 0\nCAMERA_UP 0 1 0\n"
                      "CAMERA_FORWARD 0 0 -1\nCAMERA_FOV_X 1.1\nNEW_PRIMITIVE\nELLIPSOID 0.9 0.9 0.9\nPOSITION 0 1 0\nDIELECTRIC\nIOR 1.33\nCOLOR 0.9 1 0.9\n\n"
                      "NEW_PRIMITIVE\nBOX 3 2 0.1\nPOSITION 0 1 -2\nROTATION 0 0.0871557 0 0.9961947\nMETALLIC\nCOLOR 0.9 0.9 0.9\n\n"
                      "NEW_PRIMITIVE\nELLIPSOID 0.3 0.2 0.3\nPOSITION 2 3 1\nROTATION 0.2 0.1 0 0.9746794\nEMISSION 8 7 6\n\n"
                      "NEW_PRIMITIVE\nPLANE 0 1 0\nCOLOR 0.6 0.6 0.6\n",
}


@pytest.mark.parametrize("name", sorted(EDGE_SCENES))
def test_edge_scenes_vs_oracle(rtc, oracle_lib, name):
    text = EDGE_SCENES[name]
    s = rtc.Scene(text=text, device=0)
    raw = text.encode()
    h = oracle_lib.lib.orc_scene_parse(raw, len(raw))
    a = orclib.Scene.__new__(orclib.Scene)
    a.b, a.h = oracle_lib, h
    info = np.zeros(8, np.uint32)
    oracle_lib.lib.orc_scene_info(h, info)
    (a.width, a.height, a.ray_depth, a.samples, a.nprims, a.nbvh, a.nnodes, a.nlights) = [int(v) for v in info]
    assert [s.width, s.height, s.ray_depth, s.samples, s.nprims, s.nbvh, s.nnodes, s.nlights] == info.tolist()
    # primary hits, both traversal modes
    ys, xs = np.mgrid[0:s.height, 0:s.width]
    xy = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], 1).astype(np.float32)
    o, d = s.cam.GetToRay(xy)
    want = a.intersect(o, d)
    check_hits(s.RayIntersection(o, d, rtc.TRAVERSAL_INDEX), want, id_agree=0.998)
    check_hits(s.RayIntersection(o, d, rtc.TRAVERSAL_REFTREE), want, id_agree=0.998)
    # render with the same Philox streams
    got = s.RenderSum(seed=4, sample_count=s.samples).reshape(-1, 3)
    ref, paths, rays = a.render_sum(4, 0, s.samples)
    c = s.counters()
    assert c["paths"] == paths
    # deep specular chains amplify rounding differences (total-internal-reflection thresholds, Fresnel
    # coin flips): path lengths then differ on a fraction of a percent of the paths
    chaotic = name == "specular_chain"
    assert abs(c["rays"] - rays) <= (1e-2 if chaotic else 2e-3) * rays + 2
    ok = np.isfinite(ref).all(1) & np.isfinite(got).all(1)
    rel = np.abs(got - ref).max(1) / (np.abs(ref).max(1) + 1e-3)
    assert ((rel <= 2e-3) & ok).mean() >= (0.9 if chaotic else 0.98), ((rel <= 2e-3) & ok).mean()
    img = s.Render(seed=4)
    assert img.shape == (s.height, s.width, 3)
    a.close()
    s.close()


def test_large_frame_4k(rtc, gpu_scenes):
    """BASELINE configs[4] geometry at reduced spp: 3840x2160 of the metal dragon, 2 spp (16.6 M paths)."""
    s = rtc.Scene(path=scene_path("practice5_dragon_100k_metal"), device=0)
    s.override(3840, 2160, 2)
    s.reset_counters()
    a = s.RenderSum(seed=1, sample_count=2)
    c = s.counters()
    assert c["paths"] == 3840 * 2160 * 2
    assert a.shape == (2160, 3840, 3)
    ok = np.isfinite(a)
    assert ok.mean() > 0.9999 and (a[ok] >= 0).all() and a[ok].max() > 0
    img = s.Render(seed=1)
    assert img.shape == (2160, 3840, 3) and img.max() > 0
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["practice5_dragon_10k", "practice5_dragon_100k"])
def test_rays_aimed_at_the_rendered_triangles(rtc, gpu_scenes, oracle_scenes, name):
    """Random rays almost never hit the dragon (a triangle is rendered on the plane through the origin but culled
    by the box of the real one: DESIGN.md 'Feasibility cones').  These rays are built to hit: through a point of
    the rendered triangle and a point just inside the real triangle's box, i.e. along the extreme directions the
    cones must let through.  Index traversal, reference-tree walk and oracle must name the same primitive."""
    from test_host_emulation import _aimed_rays
    tris = []
    for line in open(scene_path(name)):
        if line.startswith("TRIANGLE"):
            tris.append([float(v) for v in line.split()[1:10]])
    T = np.array(tris, np.float32).astype(np.float64).reshape(-1, 3, 3)
    rng = np.random.default_rng(77)
    o, d = _aimed_rays(rng, T, 150000)
    s, a = gpu_scenes(name), oracle_scenes(name)
    fast = s.RayIntersection(o, d, rtc.TRAVERSAL_INDEX)
    slow = s.RayIntersection(o, d, rtc.TRAVERSAL_REFTREE)
    bvh_hit = (slow[0] >= 0) & (slow[0] < s.nbvh)
    assert bvh_hit.mean() > 0.3, bvh_hit.mean()          # the rays do reach triangles
    assert (fast[0] == slow[0]).mean() >= 0.9999, (fast[0] != slow[0]).sum()
    k = 30000
    want = a.intersect(o[:k], d[:k])
    assert (fast[0][:k] == want[0]).mean() >= 0.9999, (fast[0][:k] != want[0]).sum()
    check_hits(tuple(x[:k] for x in fast), want, id_agree=0.9999)


# ------------------------------------------------------------------------------- round 2: frames, devices, probes at rate
def _devices_for_test(rtc):
    """Two devices when the box has them, else the same device twice (two replicas, the whole multi-device path
    -- sample split, per-replica render, reduce + resolve kernel -- on one GPU)."""
    return [0, 1] if rtc.device_count() >= 2 else [0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,spp", [("lights_mix", 96, 64, 16), ("practice5_dragon_10k", 128, 128, 8)])
def test_multi_device_render_equals_single_device(rtc, name, w, h, spp):
    """Scene::Render over N devices of one process (rtc_render_u8_multi): the Philox streams depend on
    (pixel, sample) only, so N sample shards summed over peer access are the 1-device frame up to float summation
    order -- float sums within 1e-4 relative, 8-bit image within 1 LSB."""
    s = rtc.Scene(path=scene_path(name), device=0)
    s.override(w, h, spp)
    one = s.RenderSum(seed=11, sample_count=spp)
    img1 = s.Render(seed=11)
    for devices in (_devices_for_test(rtc), [0, 0, 0]):
        img, total = s.RenderMulti(devices, seed=11, want_sum=True)
        ok = np.isfinite(one) & np.isfinite(total)
        assert ok.mean() > 0.999
        assert np.allclose(total[ok], one[ok], rtol=1e-4, atol=1e-5), np.abs(total[ok] - one[ok]).max()
        assert (np.abs(img.astype(int) - img1.astype(int)) <= 1).mean() > 0.999
    s.close()


@pytest.mark.gpu
def test_multi_device_cli(rtc, tmp_path):
    """RTC_DEVICES=<list> ./run.sh <scene> <out.ppm>"""
    from conftest import ROOT
    devs = ",".join(str(d) for d in _devices_for_test(rtc))
    outs = []
    for extra in ({}, {"RTC_DEVICES": devs}):
        out = tmp_path / ("cli_%d.ppm" % len(outs))
        env = dict(os.environ, RTC_SAMPLES="8", RTC_WIDTH="96", RTC_HEIGHT="64", RTC_SEED="5", **extra)
        r = subprocess.run([os.path.join(ROOT, "run.sh"), scene_path("practice5_2"), str(out)], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(np.frombuffer(out.read_bytes()[len(b"P6\n96 64\n255\n"):], np.uint8))
    assert (np.abs(outs[0].astype(int) - outs[1].astype(int)) <= 1).mean() > 0.999


@pytest.mark.gpu
def test_frames_in_flight_equal_render(rtc):
    """rtc_frame_begin / rtc_frame_end (async scene upload into the idle arena, render, resolve, image to pinned host
    memory) with two frames queued: every frame is the image rtc_render_u8 gives for its seed."""
    s = rtc.Scene(path=scene_path("practice5_dragon_10k"), device=0)
    s.override(160, 120, 4)
    want = [s.Render(seed=20 + i) for i in range(4)]
    got = [np.zeros_like(want[0]) for _ in range(4)]
    h2d = s.frame_begin(seed=20, slot=0)
    assert h2d > 1000000
    for i in range(1, 4):
        s.frame_begin(seed=20 + i, slot=i & 1)
        s.frame_end((i - 1) & 1, got[i - 1])
    s.frame_end(1, got[3])
    for a, b in zip(want, got):
        assert (np.abs(a.astype(int) - b.astype(int)) <= 1).mean() > 0.999
    with pytest.raises(rtc.RtcError):
        s.frame_end(0)          # nothing in flight
    s.close()


@pytest.mark.gpu
def test_intersect_on_device_arrays(rtc, gpu_scenes):
    """rtc_intersect_dev (rays and results stay in HBM, pooled scratch, no allocation per call) == rtc_intersect."""
    import torch
    s = gpu_scenes("practice5_dragon_10k")
    rng = np.random.default_rng(5)
    n = 200000
    o = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    pid, t, nrm, inter = s.RayIntersection(o, d)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    tid = torch.empty(n, dtype=torch.int32, device="cuda"); tt = torch.empty(n, dtype=torch.float32, device="cuda")
    tn = torch.empty((n, 3), dtype=torch.float32, device="cuda"); ti = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(2):   # the second call reuses every scratch buffer
        s.intersect_dev(n, to.data_ptr(), td.data_ptr(), tid.data_ptr(), tt.data_ptr(), tn.data_ptr(), ti.data_ptr(),
                        stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(tid.cpu().numpy(), pid)
    assert np.array_equal(tt.cpu().numpy(), t)
    assert np.array_equal(ti.cpu().numpy(), inter)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["practice5_1", "lights_mix", "rabbid", "practice5_dragon_10k", "practice5_dragon_100k"])
def test_device_built_arena_tail_is_byte_identical(rtc, name):
    """Only the head of the flattened scene is uploaded; the upper levels of the LCA table, the per-slot leaf boxes and
    (for scenes without rotated primitives) the identity rotations are rebuilt on the device.  The arena in HBM must be,
    byte for byte, the one the host would have uploaded in full -- after the load, after a synchronous re-upload and
    after asynchronous uploads into both arenas."""
    s = rtc.Scene(path=scene_path(name), device=0)
    assert s.arena_check() == 0
    full = s.stats()["device_bytes"]
    assert s.upload() == full
    assert s.arena_check() == 0
    s.override(32, 32, 1)
    for _ in range(2):
        s.upload_async()
        s.RenderSum(seed=1, sample_count=1)   # the render waits for the upload and switches to the other arena
        assert s.arena_check() == 0
    s.close()
