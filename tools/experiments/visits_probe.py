#!/usr/bin/env python
"""visits_probe.py [scene] -- host emulation (tests/host_emul/libemul.so, built by the CPU test suite): index-node visits
and primitive tests per ray of the golden secondary / random rays of a scene, and the ids against the reference."""
import ctypes as C
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
name = sys.argv[1] if len(sys.argv) > 1 else "practice5_dragon_100k"
L = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libemul.so"))
f32p = np.ctypeslib.ndpointer(np.float32, flags="C"); i32p = np.ctypeslib.ndpointer(np.int32, flags="C"); u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
L.emu_scene_load.restype = C.c_void_p; L.emu_scene_load.argtypes = [C.c_char_p]
L.emu_intersect.argtypes = [C.c_void_p, C.c_long, f32p, f32p, C.c_int, i32p, f32p, f32p, i32p, u64p]
g = np.load(os.path.join(ROOT, "tests", "golden", name + "_rays.npz"))
h = L.emu_scene_load(os.path.join(ROOT, "scenes", name + ".txt").encode())
for kind, pre in (("cam", ""), ("sec", "sec_"), ("rnd", "rnd_")):
    o, d = np.ascontiguousarray(g[kind + "_o"]), np.ascontiguousarray(g[kind + "_d"])
    n = len(o)
    pid = np.zeros(n, np.int32); t = np.zeros(n, np.float32); nrm = np.zeros((n, 3), np.float32); inter = np.zeros(n, np.int32); st = np.zeros(3, np.uint64)
    L.emu_intersect(h, n, o, d, 0, pid, t, nrm, inter, st)
    print("%s %-4s rays %6d  visits/ray %.3f  prim tests/ray %.3f  ids differ %d  fallbacks %d" % (name, kind, n, st[0] / n, st[2] / n, int((pid != g[pre + "pid"]).sum()), st[1]))
