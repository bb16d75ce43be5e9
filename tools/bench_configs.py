"""Secondary measurements of bench.py: the BASELINE.json configs that are not the headline.

    gpu_reference   configs[1] with the reference's torch formulation on the SAME GPU (the baseline that is not "GPU vs CPU")
    config2         configs[2]: kNN attack loop (ChamferkNNDist k=16) vs a PointNet++-SSG-shaped victim, B=64 per GPU, N=1024
    config3         configs[3]: GeoA3 attack iterations vs a DGCNN-shaped victim (k=20), GLOBAL B=128, N=2048, batch-sharded
    config4         configs[4]: Chamfer+Hausdorff fwd+bwd, GLOBAL B=512, N=M=16384, batch-sharded, + the final NCCL all-gather

The victims are tools/victims.py restatements "written from the public architecture" with random weights
(reference-SHAPED victims: the reference's model files do not travel to the GPU box and ship no weights).
Everything here is timed with CUDA events; multi-GPU numbers are reduced with MAX over ranks by the caller.
"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ev():
    return torch.cuda.Event(enable_timing=True)


def _median(v):
    v = sorted(v)
    return v[len(v) // 2]


def timed(fn, reps, warmup=2, flush=None):
    """median / min of `reps` CUDA-event timings of fn() (ms); optional L2 flush before every repetition."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = _ev(), _ev()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return _median(ts), min(ts)


# ---------------------------------------------------------------------------- configs[1] on the same GPU
def ref_pairwise(x, y):                                   # attack/CW/CW_utils/distance.py:15-32
    xx = torch.bmm(x, x.transpose(2, 1)); yy = torch.bmm(y, y.transpose(2, 1)); zz = torch.bmm(x, y.transpose(2, 1))
    dx = torch.arange(0, x.shape[1], device=x.device); dy = torch.arange(0, y.shape[1], device=x.device)
    rx = xx[:, dx, dx].unsqueeze(1).expand_as(zz.transpose(2, 1)); ry = yy[:, dy, dy].unsqueeze(1).expand_as(zz)
    return rx.transpose(2, 1) + ry - 2 * zz


def ref_losses(preds, gts):                               # ChamferDistance.forward + HausdorffDistance.forward (:40-70)
    out = []
    for red in (torch.mean, lambda t, dim: torch.max(t, dim=dim)[0]):
        P = ref_pairwise(gts, preds)
        out += [red(torch.min(P, 1)[0], dim=1), red(torch.min(P, 2)[0], dim=1)]
    return out


def gpu_reference(pcd, adv, ori, flush, reps=5):
    """The reference formulation (three bmm's, broadcast adds, min / mean / max, autograd; the pair matrix is built for
    Chamfer and again for Hausdorff) on this GPU at the headline workload, and its agreement with our step."""
    torch.backends.cuda.matmul.allow_tf32 = False
    a_ref = adv.detach().clone().requires_grad_(True)
    a_our = adv.detach().clone().requires_grad_(True)

    def ref():
        a_ref.grad = None
        l = ref_losses(a_ref, ori)
        torch.stack(l).sum().backward()
        return l

    def ours():
        a_our.grad = None
        c1, c2 = pcd.distance.chamfer(a_our, ori); h1, h2 = pcd.distance.hausdorff(a_our, ori)
        torch.stack([c1, c2, h1, h2]).sum().backward()
        return [c1, c2, h1, h2]

    r = [t.detach() for t in ref()]; o = [t.detach() for t in ours()]
    med, mn = timed(ref, reps, warmup=1, flush=flush)
    B, N = adv.shape[0], adv.shape[1]
    pairs = float(B) * N * N
    gmax = float(a_ref.grad.abs().max())
    out = {"ms_per_step": med, "ms_min": mn, "value": pairs / (med * 1e-3) / 1e9, "unit": "Gpair/s",
           "what": "attack/CW/CW_utils/distance.py:15-70 formulation in torch on this GPU (fp32, TF32 off), B=%d unchunked: "
                   "xx, yy, zz, P = 4 x %.1f GB per distance, built twice per step" % (B, B * N * N * 4 / 1e9),
           "hausdorff_bit_equal": bool(torch.equal(r[2], o[2]) and torch.equal(r[3], o[3])),
           "chamfer_rel_err": max(float(((a - b).abs() / b.abs()).max()) for a, b in zip(o[:2], r[:2])),
           "grad_rel_err": float((a_our.grad - a_ref.grad).abs().max()) / gmax}
    del a_ref, a_our
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------- configs[2]
class TorchChamferkNN(torch.nn.Module):
    """attack/CW/CW_utils/dist_utils.py:189-223 on the reference's formulations (bmm matrices, topk)."""

    def __init__(self, k=16, alpha=1.05, w1=5., w2=3.):
        super().__init__()
        self.k, self.alpha, self.w1, self.w2 = k, alpha, w1, w2

    def forward(self, adv, ori, weights=None, batch_avg=False):
        x, y = ori, adv
        zz = torch.bmm(x, y.transpose(2, 1))
        rx = torch.sum(x * x, -1)[:, :, None]; ry = torch.sum(y * y, -1)[:, None, :]
        P = rx + ry - 2 * zz
        chamfer = torch.min(P, 1)[0].mean(1)
        pc = adv.transpose(2, 1)
        inner = -2. * torch.matmul(pc.transpose(2, 1), pc)
        xx = torch.sum(pc ** 2, dim=1, keepdim=True)
        dist = xx + inner + xx.transpose(2, 1)
        neg_value, _ = (-dist).topk(k=self.k + 1, dim=-1)
        value = torch.mean(-(neg_value[..., 1:]), dim=-1)
        with torch.no_grad():
            thr = value.mean(-1) + self.alpha * value.std(-1)
            mask = (value > thr[:, None]).float()
        return chamfer * self.w1 + torch.mean(value * mask, dim=1) * self.w2


def config2(pcd, dev, rank, with_reference, B=64, N=1024, iters=30):
    """kNN attack (attack/KNN/KNN_attack.py loop shape): iterations/s of the device-resident loop, eager and CUDA graph;
    the same loop on the reference's torch formulations (victim's FPS loop + sort ball query included) on rank 0."""
    import victims
    synth = importlib.import_module("3dpointcloudattack_b200.synth")
    CL = pcd.cw_loop
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    ours_v = victims.PointNet2SSGVictim(pcd.pointnet2_utils.sample_and_group).to(dev).eval()
    data = synth.face_clouds(B, N, seed=77, first_sample=rank * B).to(dev)
    with torch.no_grad():
        torch.manual_seed(5); target = ours_v(data.transpose(1, 2))[0].argmax(1)
    out = {"workload": f"kNN attack vs PointNet++-SSG-shaped victim, B={B}/GPU N={N}, ChamferkNNDist(k=16), {iters} iterations (BASELINE configs[2])"}
    # eager only: the victim's farthest point sampling draws its start index with the host RNG on every forward
    # (model/pointnet2_utils.py:71), which a CUDA graph cannot replay
    atk = CL.KNNAttack(ours_v, CL.UntargetedLogitsAdvLoss(kappa=15.), pcd.dist_utils.ChamferkNNDist(knn_k=16),
                       CL.ProjectInnerClipLinf(0.1), attack_lr=1e-3, num_iter=iters)
    torch.manual_seed(9); atk.attack(data, target, seed=1)
    torch.manual_seed(9); adv_ours, _ = atk.attack(data, target, seed=1)
    out["iters_per_s_eager"] = iters / (atk.loop_ms * 1e-3)
    if with_reference:
        ref_v = victims.PointNet2SSGVictim(victims.torch_sample_and_group).to(dev).eval()
        ref_v.load_state_dict(ours_v.state_dict())
        atk = CL.KNNAttack(ref_v, CL.UntargetedLogitsAdvLoss(kappa=15.), TorchChamferkNN(16), CL.ProjectInnerClipLinf(0.1),
                           attack_lr=1e-3, num_iter=iters)
        torch.manual_seed(9); atk.attack(data, target, seed=1)
        torch.manual_seed(9); adv_ref, _ = atk.attack(data, target, seed=1)
        out["reference_torch_same_gpu_iters_per_s"] = iters / (atk.loop_ms * 1e-3)
        out["coordinates_agreeing_1e-4"] = float(((adv_ours - adv_ref).abs() < 1e-4).float().mean())
    # the same two loops with torch's DEFAULT convolution precision (cudnn.allow_tf32 = True), which is what a user of the
    # reference gets; the numbers above pin fp32 convolutions so that the two loops can be compared coordinate by coordinate
    torch.backends.cudnn.allow_tf32 = True
    try:
        atk = CL.KNNAttack(ours_v, CL.UntargetedLogitsAdvLoss(kappa=15.), pcd.dist_utils.ChamferkNNDist(knn_k=16),
                           CL.ProjectInnerClipLinf(0.1), attack_lr=1e-3, num_iter=iters)
        torch.manual_seed(9); atk.attack(data, target, seed=1)
        torch.manual_seed(9); atk.attack(data, target, seed=1)
        out["iters_per_s_eager_cudnn_tf32_default"] = iters / (atk.loop_ms * 1e-3)
        if with_reference:
            atk = CL.KNNAttack(ref_v, CL.UntargetedLogitsAdvLoss(kappa=15.), TorchChamferkNN(16), CL.ProjectInnerClipLinf(0.1),
                               attack_lr=1e-3, num_iter=iters)
            torch.manual_seed(9); atk.attack(data, target, seed=1)
            torch.manual_seed(9); atk.attack(data, target, seed=1)
            out["reference_torch_same_gpu_iters_per_s_cudnn_tf32_default"] = iters / (atk.loop_ms * 1e-3)
    finally:
        torch.backends.cudnn.allow_tf32 = False
    if with_reference:
        del ref_v
    del ours_v
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------- configs[3]
def config3(pcd, dev, rank, world, with_reference, global_B=128, N=2048, iters=4):
    """GeoA3 attack (attack/GeoA3/GeoA3_attack.py: margin loss + scale_const * (CD + 0.1 HD + curvature k=16), Adam) against
    a DGCNN-shaped victim with k=20 edge-conv graphs: STRONG scaling, the global batch of 128 is cut into 128/world per GPU.
    Also times, at this per-GPU batch, the pieces of the iteration that are this package's path."""
    import victims
    synth = importlib.import_module("3dpointcloudattack_b200.synth")
    F = pcd.functional
    LU = pcd.loss_utils
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    start, stop = pcd.sharding.shard_range(global_B, rank, world)
    B = stop - start
    torch.manual_seed(0)
    victim = victims.DGCNNVictim(lambda x, k: pcd.dgcnn.get_graph_feature(x, k=k), k=20).to(dev).eval()
    data = synth.face_clouds(B, N, seed=3, first_sample=start).to(dev)
    with torch.no_grad():
        label = victim(data.transpose(1, 2).contiguous())[0].argmax(1)
    out = {"workload": f"GeoA3 attack vs DGCNN-shaped victim (k=20), global B={global_B} -> {B}/GPU, N={N}, CD + 0.1 HD + curvature(k=16), "
                       f"{iters} timed iterations (BASELINE configs[3]); strong scaling",
           "global_batch": global_B, "per_gpu_batch": B}
    atk = pcd.geoa3_loop.GeoA3Attack(victim, classes=106, initial_const=10., lr=0.01, binary_max_steps=1, iter_max_steps=iters,
                                     hd_loss_weight=0.1, curv_loss_weight=1.0, curv_loss_knn=16, global_batch=global_B)
    atk.attack(data, label, seed=1, first_sample=start)
    atk.attack(data, label, seed=1, first_sample=start)
    out["ms_per_iteration"] = atk.loop_ms / iters
    torch.backends.cudnn.allow_tf32 = True                 # torch's default convolution precision (what a user of the reference gets)
    try:
        atk.attack(data, label, seed=1, first_sample=start)
        out["ms_per_iteration_cudnn_tf32_default"] = atk.loop_ms / iters
    finally:
        torch.backends.cudnn.allow_tf32 = False

    # the path inside one iteration, at this per-GPU batch
    ori = data.transpose(1, 2).contiguous()
    with torch.no_grad():
        normal = pcd.utility.estimate_normal(ori, 3)
        kappa_ori = LU._get_kappa_ori(ori, normal, 16)
    adv = (ori + 0.01 * torch.randn_like(ori)).requires_grad_(True)

    def geo():
        adv.grad = None
        cd = LU.chamfer_loss(adv, ori); hd = LU.hausdorff_loss(adv, ori)
        kap, _ = LU._get_kappa_adv(adv, ori, normal, 16)
        (cd + 0.1 * hd + LU.curvature_loss(adv, ori, kap, kappa_ori)).sum().backward()

    out["geometry_loss_fwd_bwd_ms"] = timed(geo, 5)[0]
    feats = [torch.randn(B, C, N, device=dev) for C in (3, 64, 64, 128)]
    out["knn_graph_ms_C3_64_64_128"] = [timed(lambda x=x: pcd.dgcnn.knn(x, 20), 5)[0] for x in feats]
    ef = []
    for x in feats:
        idx = pcd.dgcnn.knn(x, 20)
        xg = x.clone().requires_grad_(True)
        gout = torch.ones((B, 2 * x.shape[1], N, 20), device=dev)          # a dense upstream gradient, as the victim's conv hands back

        def edge():
            xg.grad = None
            torch.autograd.backward(pcd.dgcnn.get_graph_feature(xg, k=20, idx=idx), gout)
        ef.append(timed(edge, 3)[0])
        del xg, idx, gout
    out["edge_feature_fwd_bwd_ms_C3_64_64_128"] = ef
    R, mt = ctypes.c_int(0), ctypes.c_int(0)
    pcd._lib.load().pcd_nn1_query_tiling(B, N, N, ctypes.byref(R), ctypes.byref(mt))
    out["nn1_rows_per_lane"] = R.value
    ev0, ev1 = _ev(), _ev()
    ev0.record(); ev1.record(); torch.cuda.synchronize()
    sw = []
    for _ in range(5):
        F.time_next_sweep(ev0, ev1)
        F.nn1(ori.transpose(1, 2), adv.detach().transpose(1, 2), F.FORM_COL_ROW, F.NORM_MULSUM, cache=False)
        torch.cuda.synchronize()
        sw.append(ev0.elapsed_time(ev1))
    out["nn1_sweep_ms"] = _median(sw)
    out["nn1_sweep_tflops"] = 8.0 * B * N * N / (_median(sw) * 1e-3) / 1e12
    del feats, adv
    if with_reference:                                      # the reference's torch formulations, same weights, same GPU
        from geoa3_bench import ref_loss
        ref_v = victims.DGCNNVictim(victims.torch_graph_feature, k=20).to(dev).eval()
        ref_v.load_state_dict(victim.state_dict())
        for p in ref_v.parameters():
            p.requires_grad_(False)
        Bc = min(B, 32)                                     # [Bc,N,N] matrices + topk per layer
        o_c, n_c, k_c, l_c = ori[:Bc], normal[:Bc], kappa_ori[:Bc], label[:Bc]
        adv_r = (o_c + 0.01 * torch.randn_like(o_c)).requires_grad_(True)
        opt = torch.optim.Adam([adv_r], lr=1e-2)

        def it():
            logits = ref_v(adv_r)[0]
            onehot = torch.zeros_like(logits).scatter_(1, l_c.unsqueeze(1), 1.)
            cls = torch.clamp((onehot * logits).sum(1) - ((1. - onehot) * logits - onehot * 10000.).max(1)[0], min=0.)
            loss = (cls + 10.0 * ref_loss(adv_r, o_c, n_c, k_c)).sum() / Bc
            opt.zero_grad(); loss.backward(); opt.step()

        out["reference_torch_same_gpu_ms_per_iteration_at_B%d" % Bc] = timed(it, 2, warmup=1)[0]
        out["reference_batch"] = Bc
        del ref_v, adv_r
    del victim
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------- configs[4]
def config4(pcd, dev, rank, world, flush, fp32_peak, hbm_gbs, global_B=512, N=16384, reps=10):
    """Chamfer+Hausdorff forward+backward on GLOBAL B=512 clouds of 16384 points, batch-sharded (STRONG scaling), then the
    single collective of the path: all-gather of the [B] losses and the [B,N,3] perturbed clouds."""
    synth = importlib.import_module("3dpointcloudattack_b200.synth")
    F = pcd.functional
    start, stop = pcd.sharding.shard_range(global_B, rank, world)
    B = stop - start
    base = synth.face_clouds(4, N, seed=4321).to(dev)
    ori = base.repeat((B + 3) // 4, 1, 1)[:B].contiguous()
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    adv = (ori + 0.01 * torch.randn(ori.shape, device=dev, generator=gen)).requires_grad_(True)

    def step():
        adv.grad = None
        c1, c2 = pcd.distance.chamfer(adv, ori); h1, h2 = pcd.distance.hausdorff(adv, ori)
        losses = torch.stack([c1, c2, h1, h2])
        losses.sum().backward()
        return losses

    med, mn = timed(step, reps, warmup=2, flush=flush)
    ev0, ev1 = _ev(), _ev()
    ev0.record(); ev1.record(); torch.cuda.synchronize()
    sw = []
    for _ in range(5):
        flush.zero_()
        F.time_next_sweep(ev0, ev1)
        step()
        torch.cuda.synchronize()
        sw.append(ev0.elapsed_time(ev1))
    R, mt = ctypes.c_int(0), ctypes.c_int(0)
    pcd._lib.load().pcd_nn1_query_tiling(B, N, N, ctypes.byref(R), ctypes.byref(mt))
    losses = step().detach()
    gather_ms = None
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); torch.cuda.synchronize()
        g = []
        for _ in range(3):
            e0, e1 = _ev(), _ev()
            e0.record()
            all_loss = pcd.sharding.gather_batch(losses.t().contiguous(), global_B)
            all_adv = pcd.sharding.gather_batch(adv.detach(), global_B)
            e1.record(); torch.cuda.synchronize()
            g.append(e0.elapsed_time(e1))
        gather_ms = _median(g)
        assert all_loss.shape[0] == global_B and all_adv.shape[0] == global_B
        del all_adv
    sweep_ms = _median(sw)
    out = {"workload": f"chamfer+hausdorff fwd+bwd, global B={global_B} -> {B}/GPU, N=M={N} (BASELINE configs[4]); strong scaling; "
                       f"median of {reps} steps, 256 MiB L2 flush before every step",
           "global_batch": global_B, "per_gpu_batch": B, "ms_per_step": med, "ms_min": mn,
           "pairs_per_step_global": float(global_B) * N * N,
           "nn1_rows_per_lane": R.value, "sweep_ms": sweep_ms,
           "sweep_tflops": 8.0 * B * N * N / (sweep_ms * 1e-3) / 1e12,
           "sweep_frac": 8.0 * B * N * N / (sweep_ms * 1e-3) / fp32_peak if fp32_peak else None,
           "final_allgather_ms": gather_ms,
           "allgather_bytes": global_B * 4 * 4 + global_B * N * 3 * 4}
    del adv, ori
    torch.cuda.empty_cache()
    return out
