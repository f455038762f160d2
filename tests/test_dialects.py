"""The four earlier snapshots of the renderer (hw1 ray casting, hw2 Whitted, hw3 uniform-hemisphere path tracing,
hw4 cosine + light mix without triangles), each behind the same `run.sh <scene> <out.ppm>`.

Fixtures (tests/golden/hwN_<scene>.npz, tools/make_golden.py `dialects`): the 8-bit image the UNMODIFIED hwN
program (oracle/_ref/raytracing_hwN, built by oracle/Makefile from /root/reference/hwN) wrote for the scene text
stored next to it.  CPU tests pin the oracle's restatement of each snapshot against those images; GPU tests run
the CUDA path through the C-ABI against the same images and, with the same Philox streams, against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import orclib
from conftest import ROOT, golden

DETERMINISTIC = ["hw1_course_sample6", "hw2_hw2_lights"]
# more deterministic fixtures, checked on the CPU only (oracle here, the host compilation of the device code in
# test_host_emulation.py): hw2_glass = nested dielectrics, total internal reflection, a metallic wall, RAY_DEPTH 8
DETERMINISTIC_CPU = ["hw2_hw2_glass", "hw1_course_sample3", "hw1_course_sample5"]
MONTE_CARLO = ["hw3_course_sample6", "hw4_course_sample6", "hw3_course_sample4", "hw4_course_sample4", "hw3_course_sample3", "hw4_course_sample3"]


def load(fixture):
    g = golden(fixture)
    return bytes(g["text"]).decode(), int(g["dialect"]), g["u8"]


def u8_of(oracle_lib, linear, dialect):
    """Linear (H, W, 3) colours -> 8 bit the way the snapshot does (hw1: as they are; hw2+: ACES + gamma)."""
    flat = np.ascontiguousarray(linear.reshape(-1, 3), np.float32)
    out = np.zeros(flat.size, np.uint8)
    if dialect == 1:
        oracle_lib.lib.orc_flat_u8(flat.shape[0], flat, out)
    else:
        oracle_lib.lib.orc_tonemap_u8(flat.shape[0], flat, out)
    return out.reshape(linear.shape)


def check_monte_carlo_u8(a, b, fixture):
    """Two independent renders of ours (a, b) against the reference program's image `u8`, all 8-bit.  The fixture
    also holds `u8_b`, the same program at SAMPLES - 1, which shows the reference's OWN noise: it differs from ours
    on scenes without EMISSION lines (the reference leaves Primitive::emission indeterminate and then samples
    towards every box / ellipsoid as if it were a light; we define a missing emission as 0).  With
    var_ours = mse(a, b) / 2 and var_ref = mse(u8, u8_b) / 2, an unbiased render has mse(a, u8) = var_ours + var_ref:
    the measured RMSE must not exceed that prediction by more than 20 % (+ 0.3 grey levels of quantisation), and
    the mean difference must be within 4 standard errors (+ 0.05 of a grey level)."""
    g = golden(fixture)
    a, b = a.astype(np.float64), b.astype(np.float64)
    ref, ref_b = g["u8"].astype(np.float64), g["u8_b"].astype(np.float64)
    var_ours = np.mean((a - b) ** 2) / 2
    var_ref = np.mean((ref - ref_b) ** 2) / 2
    err = np.sqrt(np.mean((a - ref) ** 2))
    assert err <= 1.2 * np.sqrt(var_ours + var_ref) + 0.3, (err, var_ours, var_ref)
    diff = a - ref
    se = diff.std() / np.sqrt(diff.size)
    assert abs(diff.mean()) < 4 * se + 0.05, (diff.mean(), se)
    # and our render is not noisier than the reference's (same spp)
    assert var_ours <= 1.3 * var_ref + 0.1, (var_ours, var_ref)


# ---------------------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("fixture", DETERMINISTIC + DETERMINISTIC_CPU)
def test_oracle_deterministic_dialects_vs_reference_images(oracle_lib, fixture):
    text, dialect, want = load(fixture)
    s = orclib.Scene(oracle_lib, text=text, dialect=dialect)
    got = s.frame_u8()
    s.close()
    diff = np.abs(got.astype(int) - want.astype(int))
    if dialect == 1:
        assert diff.max() == 0
    else:  # hw2: the oracle sums in the reference's order; glm's pow / sqrt overloads may round the last bit differently
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


@pytest.mark.parametrize("fixture", MONTE_CARLO)
def test_oracle_monte_carlo_dialects_vs_reference_images(oracle_lib, fixture):
    text, dialect, want = load(fixture)
    s = orclib.Scene(oracle_lib, text=text, dialect=dialect)
    a = s.frame_u8(seed=21)
    b = s.frame_u8(seed=22)
    s.close()
    check_monte_carlo_u8(a, b, fixture)


def test_oracle_dialect_vocabulary(oracle_lib):
    """Each snapshot's reader knows only its own commands (hwN src/scene.cpp Scene::Load): hw1 ignores RAY_DEPTH /
    SAMPLES / materials, hw2 and hw3 have no light sampling, TRIANGLE exists in hw5 only."""
    text = ("DIMENSIONS 8 6\nRAY_DEPTH 3\nSAMPLES 5\nBG_COLOR 0 0 0\nCAMERA_POSITION 0 0 0\nCAMERA_RIGHT 1 0 0\nCAMERA_UP 0 1 0\n"
            "CAMERA_FORWARD 0 0 1\nCAMERA_FOV_X 1\nNEW_PRIMITIVE\nBOX 1 1 1\nPOSITION 0 0 5\nCOLOR 1 0 0\nEMISSION 2 2 2\n"
            "NEW_PRIMITIVE\nTRIANGLE 0 0 4 1 0 4 0 1 4\nCOLOR 0 1 0\n")
    shapes = {}
    for d in (1, 2, 3, 4, 5):
        s = orclib.Scene(oracle_lib, text=text, dialect=d)
        shapes[d] = (s.samples, s.ray_depth, s.nprims)
        s.close()
    assert shapes[1][0] == 0 and shapes[2][0] == 0            # no SAMPLES before hw3
    assert shapes[3][0] == shapes[4][0] == shapes[5][0] == 5
    assert shapes[2][1] == shapes[5][1] == 3


# ------------------------------------------------------------------------------------------------ CPU: the host
def test_product_reader_speaks_the_dialects(rtc, oracle_lib):
    """The product's scene reader (csrc/scene_load.cpp) and the oracle's agree on every dialect fixture: header
    values and primitive count, without a device (device = -1)."""
    for fixture in DETERMINISTIC + DETERMINISTIC_CPU + MONTE_CARLO:
        text, dialect, _ = load(fixture)
        a = orclib.Scene(oracle_lib, text=text, dialect=dialect)
        s = rtc.Scene(text=text, device=-1, dialect=dialect)
        assert s.dialect == dialect == s.lib.rtc_scene_dialect(s.h)
        assert (s.width, s.height, s.ray_depth, s.nprims) == (a.width, a.height, a.ray_depth, a.nprims)
        if dialect >= 3:
            assert s.samples == a.samples
        s.close()
        a.close()
    with pytest.raises(Exception):
        rtc.Scene(text="DIMENSIONS 4 4\n", device=-1, dialect=7)


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("fixture", DETERMINISTIC)
def test_gpu_deterministic_dialects_vs_reference_images(rtc, fixture):
    """hw1 / hw2: one deterministic frame; `Render` = Scene::Render of that snapshot."""
    text, dialect, want = load(fixture)
    s = rtc.Scene(text=text, device=0, dialect=dialect)
    got = s.Render(seed=5)
    again = s.Render(seed=6)  # no randomness in these dialects
    s.close()
    assert got.shape == want.shape
    assert np.array_equal(got, again)
    diff = np.abs(got.astype(int) - want.astype(int)).max(2)
    # device arithmetic contracts a*b+c into FMA where the reference does not: a ray that grazes a silhouette or a
    # shadow edge may fall on the other side (whole-pixel changes), everything else is within 1 LSB
    assert (diff <= 1).mean() >= 0.999, (diff > 1).sum()
    if dialect == 1:
        assert (diff == 0).mean() >= 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", DETERMINISTIC)
def test_gpu_deterministic_dialects_vs_oracle_linear(rtc, oracle_lib, fixture):
    text, dialect, _ = load(fixture)
    s = rtc.Scene(text=text, device=0, dialect=dialect)
    got = s.RenderSum(seed=1, sample_count=1).reshape(-1, 3)
    s.close()
    a = orclib.Scene(oracle_lib, text=text, dialect=dialect)
    want = a.render_sum(1, 0, 1)[0]
    a.close()
    rel = np.abs(got - want).max(1) / (np.abs(want).max(1) + 1e-3)
    assert (rel <= 1e-4).mean() >= 0.999, (rel > 1e-4).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", MONTE_CARLO)
def test_gpu_monte_carlo_dialects_vs_reference_images(rtc, oracle_lib, fixture):
    text, dialect, want = load(fixture)
    s = rtc.Scene(text=text, device=0, dialect=dialect)
    spp = s.samples
    a = u8_of(oracle_lib, s.RenderSum(seed=31, sample_count=spp) / np.float32(spp), dialect)
    b = u8_of(oracle_lib, s.RenderSum(seed=32, sample_count=spp) / np.float32(spp), dialect)
    img = s.Render(seed=31)
    s.close()
    check_monte_carlo_u8(a, b, fixture)
    # Render = accumulate + resolve on the device: the same image as tonemapping the sums on the host
    assert (np.abs(img.astype(int) - a.astype(int)) <= 1).mean() > 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,spp", [("hw3_course_sample6", 4), ("hw4_course_sample6", 4), ("hw3_course_sample4", 6), ("hw4_course_sample4", 6),
                                         ("hw3_course_sample3", 6), ("hw4_course_sample3", 6)])
def test_gpu_monte_carlo_dialects_sample_exact_vs_oracle(rtc, oracle_lib, fixture, spp):
    """Same Philox streams on both sides (the hw3 half-pixel jitter offset and its uniform-hemisphere lobe
    included): per-pixel sums agree to float rounding except where a path crosses a discontinuity differently."""
    text, dialect, _ = load(fixture)
    s = rtc.Scene(text=text, device=0, dialect=dialect)
    s.override(-1, -1, spp)
    got = s.RenderSum(seed=9, sample_begin=1, sample_count=spp).reshape(-1, 3)
    cnt = s.counters()
    w, h = s.width, s.height
    s.close()
    a = orclib.Scene(oracle_lib, text=text, dialect=dialect)
    a.override(-1, -1, spp)
    want, paths, rays = a.render_sum(9, 1, spp)
    a.close()
    assert cnt["paths"] == paths == w * h * spp
    assert abs(cnt["rays"] - rays) <= 2e-3 * rays
    ok = np.isfinite(want).all(1) & np.isfinite(got).all(1)
    rel = np.abs(got - want).max(1) / (np.abs(want).max(1) + 1e-3)
    assert ((rel <= 2e-3) & ok).mean() >= 0.99, ((rel <= 2e-3) & ok).mean()


@pytest.mark.gpu
@pytest.mark.parametrize("dialect,fixture", [(1, "hw1_course_sample6"), (2, "hw2_hw2_lights"), (4, "hw4_course_sample4")])
def test_gpu_cli_by_program_name(rtc, tmp_path, dialect, fixture):
    """hwN/run.sh runs build/raytracing_hwN: the CLI picks the dialect from the name it is called by."""
    text, d, want = load(fixture)
    assert d == dialect
    scene = tmp_path / "scene.txt"
    scene.write_text(text)
    out = tmp_path / "out.ppm"
    exe = os.path.join(ROOT, "raytracing-course_b200", "raytracing_hw%d" % dialect)
    r = subprocess.run([exe, str(scene), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    img = orclib.read_ppm(str(out))
    assert img.shape == want.shape
    diff = np.abs(img.astype(int) - want.astype(int)).max(2)
    if dialect <= 2:
        assert (diff <= 1).mean() >= 0.999
    else:
        assert np.sqrt(np.mean(diff.astype(float) ** 2)) < 12.0  # one noisy render against a converged one
