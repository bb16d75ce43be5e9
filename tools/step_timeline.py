"""Kernel timeline of one graph-replayed bench step (development tool): names, durations, gaps."""
import importlib, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
ori = synth.face_clouds(32, 4096, seed=1234).cuda(); adv = synth.perturb(ori.cpu(), 0.01, seed=99).cuda().requires_grad_(True)

def loss_fn(a, o):
    c1, c2 = pcd.distance.chamfer(a, o); h1, h2 = pcd.distance.hausdorff(a, o)
    l = torch.stack([c1, c2, h1, h2]); return l.sum(), (l,)

g = pcd.graph.GraphedLoss(loss_fn, adv, ori)
for _ in range(5): g.replay()
torch.cuda.synchronize()
FLUSH = "--flush" in sys.argv          # bench.py conditions: 256 MiB memset between the steps (cold L2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        if FLUSH:
            flush.zero_()
        g.replay()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memset" not in e.name and "FillFunctor" not in e.name],
             key=lambda e: e.time_range.start)
n = len(evs) // 3
last = None
for e in evs[n:2 * n]:
    gap = (e.time_range.start - last) if last is not None else 0.0
    print(f"{e.time_range.elapsed_us():8.1f} us  gap {gap:6.1f}  {e.name[:90]}")
    last = e.time_range.end
print("step span %.1f us" % (evs[2 * n - 1].time_range.end - evs[n].time_range.start))
