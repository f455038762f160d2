"""ctypes loaders for the CPU oracle (oracle/liboracle_rt.so) and, when built, the real
reference probe (oracle/_ref/librefprobe.so).  TEST INFRASTRUCTURE ONLY: nothing under
raytracing-course_b200/ imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_rt.so")
REFPROBE_SO = os.path.join(ORACLE_DIR, "_ref", "librefprobe.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "raytracing_hw5")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build_oracle():
    """(Re)build liboracle_rt.so if missing or stale.  gcc only, a second or two."""
    src = os.path.join(ORACLE_DIR, "rt_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle_rt.so"])
    return ORACLE_SO


class _Lib:
    """Common batch API; `pfx` is 'orc_' (oracle) or 'ref_' (real reference probe)."""

    def __init__(self, path, pfx):
        self.lib = C.CDLL(path)
        self.pfx = pfx
        L = self.lib
        g = lambda n: getattr(L, pfx + n)
        g("scene_load").restype = C.c_void_p
        g("scene_load").argtypes = [C.c_char_p]
        g("scene_free").argtypes = [C.c_void_p]
        g("scene_info").argtypes = [C.c_void_p, u32p]
        g("scene_override").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        g("scene_prims").argtypes = [C.c_void_p, i32p, f32p]
        g("scene_nodes").argtypes = [C.c_void_p, f32p, u32p]
        g("scene_root").argtypes = [C.c_void_p]
        g("scene_root").restype = C.c_uint32
        g("intersect").argtypes = [C.c_void_p, C.c_long, f32p, f32p, i32p, f32p, f32p, i32p]
        g("primitive_intersect").argtypes = [C.c_void_p, C.c_int, C.c_long, f32p, f32p, i32p, f32p, f32p, i32p]
        g("camera_rays").argtypes = [C.c_void_p, C.c_long, f32p, f32p, f32p]
        g("mix_pdf").argtypes = [C.c_void_p, C.c_long, f32p, f32p, f32p, f32p]
        g("tonemap_u8").argtypes = [C.c_long, f32p, u8p]
        if pfx == "orc_":
            L.orc_scene_parse.restype = C.c_void_p
            L.orc_scene_parse.argtypes = [C.c_char_p, C.c_long]
            L.orc_scene_load_dialect.restype = C.c_void_p
            L.orc_scene_load_dialect.argtypes = [C.c_char_p, C.c_int]
            L.orc_scene_parse_dialect.restype = C.c_void_p
            L.orc_scene_parse_dialect.argtypes = [C.c_char_p, C.c_long, C.c_int]
            L.orc_flat_u8.argtypes = [C.c_long, f32p, u8p]
            L.orc_scene_prim_order.argtypes = [C.c_void_p, i32p]
            L.orc_scene_camera.argtypes = [C.c_void_p, f32p]
            L.orc_mix_sample.argtypes = [C.c_void_p, C.c_long, f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p]
            L.orc_render_sum.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_long, C.c_long, f32p, u64p, C.c_int]
            L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
            L.orc_sort_perm_by_key.argtypes = [f32p, i32p, C.c_long, C.c_long]
            L.orc_partition_flags.argtypes = [i32p, u8p, C.c_long]
            L.orc_partition_flags.restype = C.c_long
        else:
            L.ref_mix_sample.argtypes = [C.c_void_p, C.c_long, f32p, f32p, C.c_uint32, f32p]
            L.ref_render_linear.argtypes = [C.c_void_p, C.c_long, C.c_long, f32p, C.c_int]
            L.ref_std_sort_perm.argtypes = [f32p, i32p, C.c_long, C.c_long]
            L.ref_std_partition.argtypes = [i32p, u8p, C.c_long]
            L.ref_std_partition.restype = C.c_long

    def fn(self, name):
        return getattr(self.lib, self.pfx + name)


class Scene:
    """A loaded scene on either backend with numpy-in / numpy-out batch calls."""

    def __init__(self, backend, path=None, dialect=5, text=None):
        self.b = backend
        self.dialect = dialect
        if text is not None:
            raw = text.encode() if isinstance(text, str) else bytes(text)
            self.h = backend.lib.orc_scene_parse_dialect(raw, len(raw), dialect)
        elif dialect != 5:
            self.h = backend.lib.orc_scene_load_dialect(os.fsencode(path), dialect)
        else:
            self.h = backend.fn("scene_load")(os.fsencode(path))
        if not self.h:
            raise IOError("cannot load scene " + str(path))
        info = np.zeros(8, np.uint32)
        backend.fn("scene_info")(self.h, info)
        (self.width, self.height, self.ray_depth, self.samples,
         self.nprims, self.nbvh, self.nnodes, self.nlights) = [int(v) for v in info]

    def close(self):
        if self.h:
            self.b.fn("scene_free")(self.h)
            self.h = None

    def override(self, width=-1, height=-1, samples=-1, ray_depth=-1):
        self.b.fn("scene_override")(self.h, width, height, samples, ray_depth)
        info = np.zeros(8, np.uint32)
        self.b.fn("scene_info")(self.h, info)
        self.width, self.height, self.ray_depth, self.samples = [int(v) for v in info[:4]]

    def prims(self):
        tm = np.zeros((self.nprims, 2), np.int32)
        d = np.zeros((self.nprims, 26), np.float32)
        self.b.fn("scene_prims")(self.h, tm, d)
        return tm, d

    def nodes(self):
        aabb = np.zeros((self.nnodes, 6), np.float32)
        links = np.zeros((self.nnodes, 4), np.uint32)
        if self.nnodes:
            self.b.fn("scene_nodes")(self.h, aabb, links)
        return aabb, links, int(self.b.fn("scene_root")(self.h))

    def intersect(self, o, d):
        o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
        n = o.shape[0]
        pid = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32); inter = np.zeros(n, np.int32)
        self.b.fn("intersect")(self.h, n, o, d, pid, t, nrm, inter)
        return pid, t, nrm, inter

    def primitive_intersect(self, prim, o, d):
        o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
        n = o.shape[0]
        hit = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32); inter = np.zeros(n, np.int32)
        self.b.fn("primitive_intersect")(self.h, prim, n, o, d, hit, t, nrm, inter)
        return hit, t, nrm, inter

    def camera_rays(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        n = xy.shape[0]
        o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
        self.b.fn("camera_rays")(self.h, n, xy, o, d)
        return o, d

    def mix_pdf(self, x, nrm, d):
        x = np.ascontiguousarray(x, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        out = np.zeros(x.shape[0], np.float32)
        self.b.fn("mix_pdf")(self.h, x.shape[0], x, nrm, d, out)
        return out

    def tonemap_u8(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
        out = np.zeros(rgb.shape, np.uint8)
        self.b.fn("tonemap_u8")(rgb.shape[0], rgb, out)
        return out

    # oracle only -----------------------------------------------------------
    def frame_u8(self, seed=0, nthreads=0):
        """Scene::Render of the scene's dialect: (H, W, 3) uint8 (hw1: colours as they are; else ACES + gamma)."""
        spp = 1 if self.dialect <= 2 else self.samples
        s, _, _ = self.render_sum(seed, 0, spp, nthreads=nthreads)
        mean = np.ascontiguousarray(s * np.float32(1.0 / spp))
        out = np.zeros(mean.shape, np.uint8)
        if self.dialect == 1:
            self.b.lib.orc_flat_u8(mean.shape[0], mean, out)
        else:
            self.b.lib.orc_tonemap_u8(mean.shape[0], mean, out)
        return out.reshape(self.height, self.width, 3)

    def prim_order(self):
        out = np.zeros(self.nprims, np.int32)
        self.b.lib.orc_scene_prim_order(self.h, out)
        return out

    def camera(self):
        out = np.zeros(16, np.float32)
        self.b.lib.orc_scene_camera(self.h, out)
        return out

    def mix_sample(self, x, nrm, seed, sample, bounce):
        x = np.ascontiguousarray(x, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
        out = np.zeros(x.shape, np.float32)
        self.b.lib.orc_mix_sample(self.h, x.shape[0], x, nrm, seed, sample, bounce, out)
        return out

    def render_sum(self, seed, sample_begin, sample_count, pix_begin=0, pix_end=None, nthreads=0):
        if pix_end is None:
            pix_end = self.width * self.height
        out = np.zeros((pix_end - pix_begin, 3), np.float32)
        cnt = np.zeros(2, np.uint64)
        self.b.lib.orc_render_sum(self.h, seed, sample_begin, sample_count, pix_begin, pix_end, out, cnt, nthreads)
        return out, int(cnt[0]), int(cnt[1])

    # reference probe only ----------------------------------------------------
    def ref_mix_sample(self, x, nrm, seed0):
        x = np.ascontiguousarray(x, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
        out = np.zeros(x.shape, np.float32)
        self.b.lib.ref_mix_sample(self.h, x.shape[0], x, nrm, seed0, out)
        return out

    def ref_render_linear(self, pix_begin=0, pix_end=None, nthreads=0):
        if pix_end is None:
            pix_end = self.width * self.height
        out = np.zeros((pix_end - pix_begin, 3), np.float32)
        self.b.lib.ref_render_linear(self.h, pix_begin, pix_end, out, nthreads)
        return out


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        _oracle = _Lib(build_oracle(), "orc_")
    return _oracle


def have_ref():
    return os.path.exists(REFPROBE_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = _Lib(REFPROBE_SO, "ref_")
    return _ref


def ref_dialect_bin(dialect):
    """oracle/_ref/raytracing_hwN: the unmodified reference program of that snapshot (oracle/Makefile)."""
    return os.path.join(ORACLE_DIR, "_ref", "raytracing_hw%d" % dialect)


def read_ppm(path):
    raw = open(path, "rb").read()
    magic, dims, maxv, body = raw.split(b"\n", 3)
    w, h = [int(v) for v in dims.split()]
    assert magic == b"P6" and maxv == b"255"
    return np.frombuffer(body, np.uint8, count=3 * w * h).reshape(h, w, 3)


def with_header(text, width=None, height=None, samples=None, ray_depth=None):
    """Scene text with DIMENSIONS / SAMPLES / RAY_DEPTH lines replaced (the reference programs take no overrides)."""
    out = []
    for line in text.splitlines():
        word = line.split()[0] if line.split() else ""
        if word == "DIMENSIONS" and width is not None:
            line = "DIMENSIONS %d %d" % (width, height)
        elif word == "SAMPLES" and samples is not None:
            line = "SAMPLES %d" % samples
        elif word == "RAY_DEPTH" and ray_depth is not None:
            line = "RAY_DEPTH %d" % ray_depth
        out.append(line)
    return "\n".join(out) + "\n"


def pixel_center_rays(scene, step=1):
    ys, xs = np.mgrid[0:scene.height:step, 0:scene.width:step]
    xy = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], 1).astype(np.float32)
    return scene.camera_rays(xy)
