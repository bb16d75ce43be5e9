"""Sweep time of the library selected by PCDIST_LIBRARY in both modes (development tool)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200"); F = pcd.functional
synth = importlib.import_module("3dpointcloudattack_b200.synth")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
B, N = 32, 4096
ori = synth.face_clouds(B, N, seed=1234).cuda(); adv = (ori + 0.01 * torch.randn_like(ori)).contiguous()
for mode, name in ((F.SWEEP_EXACT, "exact"), (F.SWEEP_APPROX, "approx")):
    F.force_sweep_mode(mode)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); ev1.record(); torch.cuda.synchronize()
    ts = []
    for k in range(12):
        flush.zero_()
        F.time_next_sweep(ev0, ev1)
        F.nn1(adv, ori, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1) * 1e3)
    print(os.path.basename(os.environ.get("PCDIST_LIBRARY", "libpcdist.so")), name, "sweep median %.1f us  min %.1f" % (sorted(ts)[6], min(ts)), flush=True)
