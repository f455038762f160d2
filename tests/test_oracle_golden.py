"""CPU: pins oracle/rt_oracle.c against golden vectors recorded from the UNMODIFIED reference
(tools/make_golden.py -> tests/golden/*.npz).  Integer/index results must be identical; float
results of the deterministic functions are compared BIT-FOR-BIT (the oracle is built without FMA
contraction, like the reference)."""
import numpy as np
import pytest

import orclib
from conftest import golden, scene_path

RAY_SCENES = ["practice5_1", "practice5_2", "lights_mix", "practice5_dragon_10k", "rabbid"]
# the headline scenes pinned to the reference itself (tools/make_golden.py headline): the oracle's node-by-node walk of the
# reference tree costs ~10^4 node visits per ray at 100k triangles, so the CPU suite checks a strided subset of each set
HEADLINE_SCENES = ["practice5_dragon_100k", "practice5_dragon_100k_glass", "practice5_dragon_100k_metal"]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_philox_known_answers(oracle_lib):
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        out = np.zeros(4, np.uint32)
        oracle_lib.lib.orc_philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32), out)
        assert out.tolist() == want


def test_libstdcxx_sort_and_partition_order(oracle_lib):
    g = golden("libstdcxx_order")
    i = 0
    while "sort%d_key" % i in g:
        key = g["sort%d_key" % i]
        perm = np.arange(len(key), dtype=np.int32)
        oracle_lib.lib.orc_sort_perm_by_key(key, perm, 0, len(key))
        assert np.array_equal(perm, g["sort%d_perm" % i]), "std::sort order, case %d" % i
        pred = g["part%d_pred" % i]
        perm2 = np.arange(len(pred), dtype=np.int32)
        cut = oracle_lib.lib.orc_partition_flags(perm2, pred, len(pred))
        assert cut == int(g["part%d_cut" % i][0])
        assert np.array_equal(perm2, g["part%d_perm" % i]), "std::partition order, case %d" % i
        i += 1
    assert i >= 8


@pytest.mark.parametrize("name", RAY_SCENES)
def test_scene_structure(oracle_scenes, name):
    g = golden(name + "_rays")
    s = oracle_scenes(name)
    assert [s.width, s.height, s.ray_depth, s.samples, s.nprims, s.nbvh, s.nnodes, s.nlights] == g["info"].tolist()
    tm, data = s.prims()
    assert np.array_equal(tm, g["prim_type_material"])
    if name.startswith("practice5_dragon") or name == "rabbid":  # (plane POSITION is uninitialised memory in the reference otherwise)
        assert np.bitwise_xor.reduce(data.view(np.uint32).ravel()) == g["prim_data_crc"][0]
    aabb, links, root = s.nodes()
    crc = np.bitwise_xor.reduce((links.ravel().astype(np.uint64) * np.arange(1, links.size + 1, dtype=np.uint64)) & np.uint64(0xFFFFFFFF))
    assert crc == g["node_links_crc"][0]
    assert root == int(g["root"][0])
    assert np.array_equal(aabb.astype(np.float64).sum(0), g["node_aabb_sum"])


@pytest.mark.parametrize("name", RAY_SCENES)
def test_camera_rays_bit_exact(oracle_scenes, name):
    g = golden(name + "_rays")
    o, d = oracle_scenes(name).camera_rays(g["xy"])
    assert np.array_equal(bits(o), bits(g["cam_o"]))
    assert np.array_equal(bits(d), bits(g["cam_d"]))


@pytest.mark.parametrize("name", RAY_SCENES)
@pytest.mark.parametrize("kind", ["cam", "sec", "rnd"])
def test_ray_intersection_bit_exact(oracle_scenes, name, kind):
    g = golden(name + "_rays")
    pre = {"cam": "", "sec": "sec_", "rnd": "rnd_"}[kind]
    o, d = g[kind + "_o"], g[kind + "_d"]
    pid, t, nrm, inter = oracle_scenes(name).intersect(o, d)
    assert np.array_equal(pid, g[pre + "pid"])
    assert np.array_equal(bits(t), bits(g[pre + "t"]))
    assert np.array_equal(bits(nrm), bits(g[pre + "nrm"]))
    assert np.array_equal(inter, g[pre + "inter"])


@pytest.mark.parametrize("name", RAY_SCENES)
def test_mix_pdf_bit_exact(oracle_scenes, name):
    g = golden(name + "_rays")
    pdf = oracle_scenes(name).mix_pdf(g["pdf_x"], g["pdf_n"], g["pdf_d"])
    assert np.array_equal(bits(pdf), bits(g["pdf"]))


@pytest.mark.parametrize("name", RAY_SCENES)
def test_mix_pdf_towards_lights_bit_exact(oracle_scenes, name):
    """Directions drawn by the reference's own Distribution::Sample: half of them hit a light, many graze it."""
    g = golden(name + "_rays")
    pdf = oracle_scenes(name).mix_pdf(g["pdf_x"], g["pdf_n"], g["lpdf_d"])
    assert np.array_equal(bits(pdf), bits(g["lpdf"]))


def test_primitive_intersect_bit_exact(oracle_scenes):
    g = golden("primitive_intersect")
    n = 0
    for name in ("practice5_2", "lights_mix"):
        s = oracle_scenes(name)
        for prim in range(s.nprims):
            key = "%s_%d" % (name, prim)
            hit, t, nrm, inter = s.primitive_intersect(prim, g[key + "_o"], g[key + "_d"])
            assert np.array_equal(hit, g[key + "_hit"]), key
            assert np.array_equal(bits(t), bits(g[key + "_t"])), key
            assert np.array_equal(bits(nrm), bits(g[key + "_nrm"])), key
            assert np.array_equal(inter, g[key + "_inter"]), key
            n += 1
    assert n == 13


def test_tonemap_bit_exact(oracle_scenes):
    g = golden("tonemap")
    assert np.array_equal(oracle_scenes("practice5_1").tonemap_u8(g["rgb"]), g["u8"])


def squash(x):
    return x / (1.0 + x)


@pytest.mark.parametrize("name,w,h,spp", [("practice5_1", 64, 48, 256), ("practice5_2", 64, 48, 1024), ("lights_mix", 48, 32, 1024),
                                          ("rabbid", 88, 88, 256)])
def test_render_statistically_matches_reference(oracle_lib, name, w, h, spp):
    """The oracle's Philox-driven integrator against the reference's own render (minstd streams):
    same expectation, independent noise.  RMSE against the reference must not exceed the RMSE
    between two independent oracle renders by more than 20 %, and the image means must agree
    within 4 standard errors."""
    g = golden("%s_render_%dx%d_%dspp" % (name, w, h, spp))
    ref = g["mean"].astype(np.float64)
    s = orclib.Scene(oracle_lib, scene_path(name))
    s.override(w, h, spp)
    a = s.render_sum(11, 0, spp)[0].reshape(h, w, 3).astype(np.float64) / spp
    b = s.render_sum(12, 0, spp)[0].reshape(h, w, 3).astype(np.float64) / spp
    s.close()
    ok = np.isfinite(ref) & np.isfinite(a) & np.isfinite(b)
    assert ok.mean() > 0.999
    sa, sb, sr = squash(a[ok]), squash(b[ok]), squash(ref[ok])
    noise = np.sqrt(np.mean((sa - sb) ** 2))
    err = np.sqrt(np.mean((sa - sr) ** 2))
    assert err <= 1.2 * noise + 1e-4, (err, noise)
    # bias: per-pixel differences have zero mean
    diff = sa - sr
    se = diff.std() / np.sqrt(diff.size) + 1e-12
    assert abs(diff.mean()) < 4 * se + 2e-4, (diff.mean(), se)


# values the reference leaves INDETERMINATE (a missing third argument of POSITION / BOX is never
# written: glm vectors are not zero-initialised); we define them as 0
PARSER_INDETERMINATE = {3: [(0, 8), (1, 16)]}


def test_parser_quirks_match_reference(oracle_lib):
    """Scene::Load on unusual inputs, recorded from the reference (tools/make_golden.py parser_fixture)."""
    g = golden("parser_quirks")
    i = 0
    while "text%d" % i in g:
        raw = g["text%d" % i].tobytes()
        h = oracle_lib.lib.orc_scene_parse(raw, len(raw))
        info = np.zeros(8, np.uint32)
        oracle_lib.lib.orc_scene_info(h, info)
        assert info.tolist() == g["info%d" % i].tolist(), i
        n = int(info[4])
        tm = np.zeros((n, 2), np.int32)
        d = np.zeros((n, 26), np.float32)
        oracle_lib.lib.orc_scene_prims(h, tm, d)
        assert np.array_equal(tm, g["tm%d" % i]), i
        want = g["data%d" % i].copy()
        for r, c in PARSER_INDETERMINATE.get(i, []):
            want[r, c] = d[r, c]
        assert np.array_equal(d, want), (i, np.argwhere(d != want).tolist())
        oracle_lib.lib.orc_scene_free(h)
        i += 1
    assert i == 4


@pytest.mark.parametrize("name", HEADLINE_SCENES)
def test_headline_scene_structure_and_rays_bit_exact(oracle_scenes, name):
    """The 100k dragons against golden vectors of the compiled reference: tree (node count, links checksum, box sums),
    camera rays, and primitive id / t / normal of primary, first-bounce secondary and random rays, bit for bit."""
    g = golden(name + "_rays")
    s = oracle_scenes(name)
    info = g["info"]
    assert [s.width, s.height, s.ray_depth, s.samples, s.nprims, s.nbvh, s.nnodes, s.nlights] == info.tolist()
    aabb, links, root = s.nodes()
    crc = np.bitwise_xor.reduce((links.ravel().astype(np.uint64) * np.arange(1, links.size + 1, dtype=np.uint64)) & np.uint64(0xFFFFFFFF))
    assert crc == g["node_links_crc"][0] and root == g["root"][0]
    assert np.array_equal(aabb.astype(np.float64).sum(0), g["node_aabb_sum"])
    o, d = s.camera_rays(g["xy"])
    assert np.array_equal(bits(o), bits(g["cam_o"])) and np.array_equal(bits(d), bits(g["cam_d"]))
    step = 16 if name == "practice5_dragon_100k" else 8
    for kind, pre in (("cam", ""), ("sec", "sec_"), ("rnd", "rnd_")):
        sel = slice(0, None, step)
        pid, t, nrm, inter = s.intersect(g[kind + "_o"][sel], g[kind + "_d"][sel])
        assert np.array_equal(pid, g[pre + "pid"][sel]), (kind, (pid != g[pre + "pid"][sel]).sum())
        assert np.array_equal(bits(t), bits(g[pre + "t"][sel]))
        assert np.array_equal(bits(nrm), bits(g[pre + "nrm"][sel]))
        assert np.array_equal(inter, g[pre + "inter"][sel])
