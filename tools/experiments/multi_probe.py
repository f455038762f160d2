#!/usr/bin/env python
"""multi_probe.py N -- Scene::Render of practice5_dragon_100k through rtc_render_u8_multi on 1..N devices of this process:
wall time per frame (scene resident, host image out) and Mpaths/s."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import raytracing_course_b200 as rtc
n = int(sys.argv[1]) if len(sys.argv) > 1 else rtc.device_count()
s = rtc.Scene(path=os.path.join(ROOT, "scenes", "practice5_dragon_100k.txt"), device=0)
paths = s.width * s.height * s.samples
out = {}
k = 1
while k <= n:
    devs = list(range(k))
    for i in range(3):
        s.RenderMulti(devs, seed=i)
    t0 = time.perf_counter()
    reps = 10
    for i in range(reps):
        s.RenderMulti(devs, seed=10 + i)
    dt = (time.perf_counter() - t0) / reps
    out[k] = {"ms_per_frame": dt * 1e3, "mpaths_per_s": paths / dt / 1e6}
    k *= 2
print(json.dumps({"workload": "practice5_dragon_100k", "api": "rtc_render_u8_multi (one process, peer-access reduce + resolve)", "devices": out}))
s.close()
