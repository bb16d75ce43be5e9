"""Timeline of one PipelinedLoss replay (development tool): every device activity with start offset, duration, stream."""
import importlib, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8, 8, 8, 8]
ori_h = synth.face_clouds(32, 4096, seed=1234).pin_memory(); adv_h = synth.perturb(ori_h, 0.01, seed=99).pin_memory()
def loss_fn(a, o):
    c1, c2 = pcd.distance.chamfer(a, o); h1, h2 = pcd.distance.hausdorff(a, o)
    l = torch.stack([c1, c2, h1, h2]); return l.sum(), (l,)
piped = pcd.graph.PipelinedLoss(loss_fn, adv_h, ori_h, slice_sizes=sizes)
for _ in range(5): piped.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): piped.replay()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
n = len(evs) // 3
sel = evs[n:2 * n]
t0 = sel[0].time_range.start
for e in sel:
    print(f"{e.time_range.start - t0:8.1f} +{e.time_range.elapsed_us():7.1f} us  {e.name[:70]}")
print("span %.1f us" % (max(e.time_range.end for e in sel) - t0))
