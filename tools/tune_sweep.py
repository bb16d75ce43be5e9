"""GPU tuning / probing helper (not part of the product or the tests).

  python tools/tune_sweep.py            # sweep-kernel time for every (R, MT) at BASELINE config 2
  python tools/tune_sweep.py --probe    # does torch's GPU path (cuBLAS K=3) match the oracle bit for bit?
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
F = pcd.functional


def time_sweep(B, N, M, reps=10, form=F.FORM_SUM_FIRST, norm=F.NORM_FMA):
    lib = pcd._lib.load()
    ori = synth.face_clouds(min(B, 4), N, seed=1).cuda().repeat((B + 3) // 4, 1, 1)[:B].contiguous()
    adv = (ori + 0.01 * torch.randn_like(ori))[:, :M].contiguous() if M <= N else torch.randn(B, M, 3, device="cuda")
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    t0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    t1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for e in e0 + e1:
        e.record()
    torch.cuda.synchronize()
    for _ in range(3):
        F.nn1(ori, adv, form, norm, cache=False)
    for k in range(reps):
        F.time_next_sweep(e0[k], e1[k])
        t0[k].record()
        F.nn1(ori, adv, form, norm, cache=False)
        t1[k].record()
    torch.cuda.synchronize()
    sw = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    tot = sorted(a.elapsed_time(b) for a, b in zip(t0, t1))
    return sw[len(sw) // 2], tot[len(tot) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe", action="store_true")
    ap.add_argument("--scaling", action="store_true")
    args = ap.parse_args()
    print("device", torch.cuda.get_device_name(0), "fp32 peak TFLOP/s %.1f" % (F.fp32_peak_flops(2048) / 1e12))
    if args.probe:
        import numpy as np
        from oracle import pcd_oracle as O
        from oracle import ref_torch_port as RP
        torch.backends.cuda.matmul.allow_tf32 = False
        ori = synth.face_clouds(2, 1024, seed=1); adv = synth.perturb(ori, 0.01, seed=2)
        P = RP.batch_pairwise_dist(ori.cuda(), adv.cuda()).cpu().numpy()
        Po = O.batch_pairwise_dist(ori.numpy(), adv.numpy())
        print("torch-GPU batch_pairwise_dist == oracle (bitwise fraction): %.6f  max|diff| %.3e" % ((P == Po).mean(), np.abs(P - Po).max()))
        print("  argmin agreement rows %.6f cols %.6f" % ((P.argmin(2) == Po.argmin(2)).mean(), (P.argmin(1) == Po.argmin(1)).mean()))
        a = ori.permute(0, 2, 1).contiguous().cuda(); b = adv.permute(0, 2, 1).contiguous().cuda()
        Mg = torch.cdist(a.permute(0, 2, 1), b.permute(0, 2, 1)).cpu().numpy()
        Mo = O.dis_pairwise_distances(a.cpu().numpy(), b.cpu().numpy())
        print("torch-GPU cdist == oracle: %.6f  max|diff| %.3e" % ((Mg == Mo).mean(), np.abs(Mg - Mo).max()))
        return
    if args.scaling:
        for R in (8, 16):
            F.force_tiling(R, 0)
            for B in (8, 16, 32, 64, 128, 256):
                sw, tot = time_sweep(B, 4096, 4096)
                print(f"R={R} B={B:4d} N=M=4096: sweep {sw*1e3:8.1f} us   {B*4096*4096/sw/1e9:6.2f} Tpair/s")
        return
    for (B, N, M) in [(32, 4096, 4096), (1, 4096, 4096), (64, 1024, 1024), (8, 16384, 16384)]:
        pairs = B * N * M
        for R in (2, 4, 8, 16):
            for MT in (128, 256):
                F.force_tiling(R, MT)
                sw, tot = time_sweep(B, N, M)
                print(f"B={B} N={N} M={M} R={R} MT={MT}: sweep {sw*1e3:8.1f} us  fwd total {tot*1e3:8.1f} us  "
                      f"{pairs/sw/1e9:8.2f} Tpair/s  {8*pairs/sw/1e9:6.1f} TFLOP/s")
        os.environ.pop("PCD_SWEEP_R"); os.environ.pop("PCD_SWEEP_MT")
        sw, tot = time_sweep(B, N, M)
        print(f"B={B} N={N} M={M} heuristic: sweep {sw*1e3:8.1f} us  fwd total {tot*1e3:8.1f} us  {pairs/sw/1e9:8.2f} Tpair/s")


if __name__ == "__main__":
    main()
