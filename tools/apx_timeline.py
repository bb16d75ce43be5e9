"""Kernel durations of one NN-1 forward in both sweep modes (development tool)."""
import importlib, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200"); F = pcd.functional
synth = importlib.import_module("3dpointcloudattack_b200.synth")
B, N = 32, 4096
ori = synth.face_clouds(B, N, seed=1234).cuda(); adv = (ori + 0.01 * torch.randn_like(ori)).contiguous()
for mode in (F.SWEEP_EXACT, F.SWEEP_APPROX):
    F.force_sweep_mode(mode)
    for _ in range(3): F.nn1(adv, ori, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        F.nn1(adv, ori, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)
        torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print("mode", mode)
    for e in evs:
        print(f"  {e.time_range.start - t0:8.1f} +{e.time_range.elapsed_us():7.1f} us  {e.name[:80]}")
