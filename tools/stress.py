"""Large-shape smoke / consistency run (development tool): sizes beyond the test-suite, checked through
size-independent properties (role swap, idempotence, self-neighbour, agreement between code paths)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
F = pcd.functional

def main():
    t0 = time.time()
    # NN-1 at the full BASELINE configs[4] per-GPU shard and at a large batch
    for (B, N) in [(64, 16384), (512, 4096)]:
        ori = synth.face_clouds(8, N, seed=7).cuda().repeat(B // 8, 1, 1).contiguous()
        adv = (ori + 0.01 * torch.randn_like(ori))
        a = F.nn1(ori, adv, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)
        b = F.nn1(adv, ori, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)
        assert torch.equal(a.row_min, b.col_min) and torch.equal(a.col_arg, b.row_arg) and torch.equal(a.row_arg, b.col_arg)
        # brute-force check of 64 random rows against torch on the same device (fp64 distances, argmin must agree
        # wherever the fp32 minimum is unique by a safe margin)
        ii = torch.randint(0, N, (64,), device="cuda")
        d64 = ((ori[0, ii].double()[:, None] - adv[0].double()[None]) ** 2).sum(-1)
        top2 = d64.topk(2, largest=False)
        safe = (top2.values[:, 1] - top2.values[:, 0]) > 1e-6
        assert torch.equal(a.row_arg[0, ii][safe].long(), top2.indices[:, 0][safe])
        print(f"nn1 B={B} N={N}: ok ({int(safe.sum())}/64 rows checked against fp64 brute force)", flush=True)
    # k-NN xyz: collect pipeline vs warp-per-row select on a big batch
    x = torch.rand(512, 2048, 3, device="cuda")
    d1, i1 = F.knn(x, x, 20)
    F.force_knn_strategy(F.KNN_BOUND_SELECT)
    d2, i2 = F.knn(x, x, 20)
    F.force_knn_strategy(F.KNN_AUTO)
    assert torch.equal(i1, i2) and torch.equal(d1, d2)
    # (expansion-form self distances are rounding noise, ~1e-7: a neighbour closer than ~6e-4 can legitimately win)
    self_first = float((i1[:, :, 0].long() == torch.arange(2048, device="cuda").expand(512, -1)).float().mean())
    assert self_first > 0.9999
    print(f"knn xyz B=512 N=2048 K=20: collect pipeline == select kernel, self first on {self_first:.6f} of the rows", flush=True)
    # feature k-NN: self first, distances ascending
    f = torch.randn(64, 4096, 64, device="cuda")
    d, i = F.knn(f, f, 20)
    assert float((i[:, :, 0].long() == torch.arange(4096, device="cuda").expand(64, -1)).float().mean()) > 0.9999
    assert bool((d[:, :, 1:] >= d[:, :, :-1]).all())
    print("knn C=64 B=64 N=4096 K=20: ok", flush=True)
    # edge features: forward == torch gather, backward == autograd of the torch formulation
    xf = torch.randn(64, 64, 2048, device="cuda", requires_grad=True)
    out = pcd.dgcnn.get_graph_feature(xf, k=20, idx=i1[:64].long())
    xr = xf.detach().clone().requires_grad_(True)
    idx = i1[:64].long()
    nb = torch.gather(xr[:, :, None, :].expand(-1, -1, 2048, -1), 3, idx[:, None].expand(-1, 64, -1, -1))
    ref = torch.cat((nb - xr[:, :, :, None], xr[:, :, :, None].expand(-1, -1, -1, 20)), 1)
    assert torch.equal(out, ref)
    g = torch.randn_like(out)
    out.backward(g); ref.backward(g)
    assert float((xf.grad - xr.grad).abs().max() / xr.grad.abs().max()) < 1e-5
    print("edge features B=64 C=64 N=2048 k=20: forward bit-equal, backward 1e-5", flush=True)
    # FPS against the torch loop on CPU-order arithmetic is covered by the tests; here: distinct indices, start kept
    xyz = torch.rand(256, 4096, 3, device="cuda")
    s = F.farthest_point_sample(xyz, 1024)
    assert bool((s[:, 0] == 0).all()) and all(len(set(r.tolist())) == 1024 for r in s[:4])
    print("fps B=256 N=4096 npoint=1024: ok", flush=True)
    # ---- round-2 kernels at BASELINE sizes, through size-independent properties / the torch chain on the same GPU
    B, N, K = 128, 2048, 16
    pc = synth.face_clouds(8, N, seed=3).cuda().repeat(B // 8, 1, 1).contiguous()
    pc = pc + 1e-3 * torch.randn_like(pc)
    _, idx = F.knn(pc, pc, K + 1)
    normal, evecs, evals = F.local_frames(pc, idx, skip_first=True, normals=True, frames=True)
    nl = normal.norm(dim=2)                     # unit, or exactly zero where the reference's sign(<n, sum of offsets>) is sign(0)
    zero = nl == 0
    assert float((nl[~zero] - 1).abs().max()) < 1e-5 and float(zero.float().mean()) < 1e-3
    gram = evecs @ evecs.transpose(2, 3)
    assert float((gram - torch.eye(3, device="cuda")).abs().max()) < 1e-5
    assert bool((evals[:, :, 1:] >= evals[:, :, :-1]).all()) and float(evals.min()) > -1e-6
    assert float((((normal * evecs[:, :, 0]).sum(2).abs() - 1).abs())[~zero].max()) < 1e-5      # the smallest eigenvector, up to sign
    pcg = pc.clone().requires_grad_(True)
    kap = F.kappa(pcg, normal, idx, skip_first=True)
    assert float(kap.detach().min()) >= 0.0 and float(kap.detach().max()) <= 1.0 + 1e-6
    kap.sum().backward()
    assert bool(torch.isfinite(pcg.grad).all())
    print(f"local frames + kappa B={B} N={N} K={K}: unit normals, orthonormal frames, ascending eigenvalues, kappa in [0,1]", flush=True)
    # k-NN outlier loss against its nine-op torch chain
    d, _ = F.knn(pc[:64, :1024], pc[:64, :1024], 6)
    value = d[..., 1:].mean(-1)
    thr = value.mean(1) + 1.05 * value.std(1)
    ref = (value * (value > thr[:, None]).float()).mean(1)
    ours = pcd.dist_utils.KNNDist(k=5, alpha=1.05)(pc[:64, :1024], batch_avg=False)
    assert float((ours - ref).abs().max() / ref.abs().max()) < 1e-5
    print("kNN outlier loss B=64 N=1024 k=5: 1e-5 of the torch chain", flush=True)
    # clip epilogues against the torch chains on this GPU (bit-equal for the per-point clips)
    ori_cf = pc.transpose(1, 2).contiguous()
    adv_cf = ori_cf + 0.03 * torch.randn_like(ori_cf)
    dcl = adv_cf - ori_cf
    t_linf = ori_cf + dcl * torch.clamp(0.03 / (torch.sum(dcl ** 2, dim=1) ** 0.5 + 1e-9), max=1.)[:, None, :]
    assert torch.equal(F.clip_points_(adv_cf.clone(), ori_cf, 0.03), t_linf)
    ln = (dcl ** 2).sum(1, keepdim=True).sqrt()
    t_lp = torch.where(ln < 0.02, dcl, torch.where(ln > 1e-6, dcl / ln.expand_as(dcl) * 0.02, torch.zeros_like(dcl)))
    assert torch.equal(F.lp_clip(dcl, 0.02), t_lp)
    t_l2 = ori_cf + dcl * torch.clamp(0.5 / (torch.sum(dcl ** 2, dim=[1, 2]) ** 0.5 + 1e-9), max=1.)[:, None, None]
    assert float((F.clip_points_(adv_cf.clone(), ori_cf, 0.5, F.CLIP_L2) - t_l2).abs().max()) < 1e-6
    print(f"clip epilogues B={B} N={N}: ClipPointsLinf and lp_clip bit-equal to the torch chains, ClipPointsL2 1e-6", flush=True)
    # reproducible edge-feature backward at the DGCNN layer-2 size
    xf = torch.randn(128, 64, 2048, device="cuda", requires_grad=True)
    gidx = pcd.dgcnn.knn(pc.transpose(1, 2).contiguous(), 20)
    out = pcd.dgcnn.get_graph_feature(xf, k=20, idx=gidx)
    g = torch.randn_like(out)
    ga = torch.autograd.grad(out, xf, g, retain_graph=True)[0]
    F.deterministic_edge_backward(True)
    g1 = torch.autograd.grad(out, xf, g, retain_graph=True)[0]
    g2 = torch.autograd.grad(out, xf, g, retain_graph=True)[0]
    F.deterministic_edge_backward(False)
    assert torch.equal(g1, g2) and float((g1 - ga).abs().max() / ga.abs().max()) < 1e-5
    print("edge-feature backward B=128 C=64 N=2048 k=20: gather form bit-reproducible, 1e-5 of the atomics form", flush=True)
    del out, g, ga, g1, g2, xf
    torch.cuda.synchronize()
    print("stress ok in %.1f s, peak memory %.1f GB" % (time.time() - t0, torch.cuda.max_memory_allocated() / 1e9))

if __name__ == "__main__":
    main()
