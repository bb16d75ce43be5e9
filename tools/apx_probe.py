"""Exact vs approximate sweep (development tool): sweep time, step time, flagged fraction, identical outputs."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200"); F = pcd.functional
synth = importlib.import_module("3dpointcloudattack_b200.synth")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (B, N, sigma) in [(32, 4096, 0.01), (32, 4096, 1e-7), (8, 4096, 0.01), (64, 16384, 0.01), (128, 2048, 0.01), (32, 4096, 1.0)]:
    ori = synth.face_clouds(min(B, 32), N, seed=1234).cuda().repeat((B + 31) // 32, 1, 1)[:B].contiguous()
    adv = (ori + sigma * torch.randn_like(ori)).contiguous()
    res = {}
    for mode, name in ((F.SWEEP_EXACT, "exact"), (F.SWEEP_APPROX, "approx")):
        F.force_sweep_mode(mode)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); ev1.record(); torch.cuda.synchronize()      # create the underlying CUDA events
        ts, tt = [], []
        for k in range(8):
            flush.zero_()
            F.time_next_sweep(ev0, ev1)
            r = F.nn1(adv, ori, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1) * 1e3)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = F.nn1(adv, ori, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False); e1.record(); torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1) * 1e3)
        res[name] = (sorted(ts)[len(ts) // 2], sorted(tt)[len(tt) // 2], r)
    F.force_sweep_mode(F.SWEEP_AUTO)
    a, b = res["exact"][2], res["approx"][2]
    same = all(torch.equal(x, y) for x, y in ((a.row_min, b.row_min), (a.row_arg, b.row_arg), (a.col_min, b.col_min), (a.col_arg, b.col_arg),
                                              (a.row_sum, b.row_sum), (a.col_max, b.col_max)))
    print(f"B={B} N={N} sigma={sigma}: sweep exact {res['exact'][0]:8.1f} us  approx {res['approx'][0]:8.1f} us ({res['exact'][0] / res['approx'][0]:.3f}x)   "
          f"forward chain exact {res['exact'][1]:8.1f}  approx {res['approx'][1]:8.1f} us   identical outputs: {same}", flush=True)
