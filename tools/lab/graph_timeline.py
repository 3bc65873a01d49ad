"""Timeline of ONE replay of the captured UNet step graph, from CUPTI activity records (torch.profiler/Kineto): start and
duration of every kernel with programmatic-dependent-launch overlap visible - the per-launch numbers of ncu are serialised and
cold, the eager CUDA-event numbers include launch gaps.  Prints per-kernel (start, duration, gap to the previous kernel's end)
and per-family sums of 'exposed' time (the part of a kernel not overlapped by its predecessor)."""
import os
import sys
import re
from collections import defaultdict
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
from tsd_b200.api import Context, Diffusion  # noqa: E402

ctx = Context(0)
m = Diffusion(ctx, 64, 64, max_batch=1)
m.init_random(1234)
rng = np.random.default_rng(0)
x = rng.standard_normal((4, 64, 64), dtype=np.float32)
cx = rng.standard_normal((77, 768), dtype=np.float32)
t = np.concatenate([np.ones(160, np.float32), np.zeros(160, np.float32)])
for _ in range(5):
    m.forward(x, cx, t)
ctx.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        m.forward(x, cx, t)
    ctx.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
print(len(ev), "kernel records")
# split into replays: a gap > 30 us between kernels separates forwards
runs, cur = [], []
for e in ev:
    if cur and e.time_range.start - cur[-1].time_range.end > 30:
        runs.append(cur)
        cur = []
    cur.append(e)
runs.append(cur)
run = runs[-2] if len(runs) >= 2 else runs[-1]
t0 = run[0].time_range.start
fam = defaultdict(lambda: [0, 0.0, 0.0])
prev_end = t0
print(f"replay with {len(run)} kernels, span {run[-1].time_range.end - t0:.1f} us")
for i, e in enumerate(run):
    name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("tsd::", "")
    name = re.sub(r"\(.*", "", name)
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    exposed = e.time_range.end - max(prev_end, e.time_range.start)
    gap = e.time_range.start - prev_end
    prev_end = max(prev_end, e.time_range.end)
    k = name[:40]
    fam[k][0] += 1
    fam[k][1] += d
    fam[k][2] += max(exposed, 0.0) + max(gap, 0.0)
    if os.environ.get("VERBOSE"):
        print(f"{i:3d} {s:8.1f} dur {d:6.1f} exposed {max(exposed, 0.0) + max(gap, 0.0):6.1f} gap {gap:6.1f}  {name[:60]}")
print("kernel                                     n   sum dur   sum exposed+gap (us)")
for k, (n, d, x_) in sorted(fam.items(), key=lambda kv: -kv[1][2]):
    print(f"{k:42s} {n:3d} {d:9.1f} {x_:9.1f}")
