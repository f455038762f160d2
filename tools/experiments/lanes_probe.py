#!/usr/bin/env python
"""lanes_probe.py [scene] -- one frame with the counting variant of k_traverse: per scheduler kind the warp iterations
and the lanes that took part (of 32), visits / tests per traversed ray."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import raytracing_course_b200 as rtc
name = sys.argv[1] if len(sys.argv) > 1 else "practice5_dragon_100k"
s = rtc.Scene(path=os.path.join(ROOT, "scenes", name + ".txt"), device=0)
acc = torch.zeros(s.width * s.height * 3, dtype=torch.float32, device="cuda")
s.render_accumulate(acc.data_ptr(), seed=1, sample_begin=0, sample_count=s.samples)
torch.cuda.synchronize()
s.reset_counters()
s.set_profiling(False, True)
s.render_accumulate(acc.data_ptr(), seed=2, sample_begin=0, sample_count=s.samples)
torch.cuda.synchronize()
c = s.counters()
l = s.traverse_lanes()
tr = max(c["traversed_rays"], 1)
out = {"scene": name, "traversed_rays": tr, "visits_per_ray": c["index_node_visits"] / tr, "tests_per_ray": c["prim_tests"] / tr,
       "lanes": l, "warp_iterations_per_ray": sum(v["iterations"] for v in l.values()) / tr}
print(json.dumps(out))
