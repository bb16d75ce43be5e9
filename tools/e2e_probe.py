"""e2e probe (development tool): monolithic graph step with host copies vs. PipelinedLoss (one graph, two branches)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
B, N = 32, 4096
ori_h = synth.face_clouds(B, N, seed=1234).pin_memory(); adv_h = synth.perturb(ori_h, 0.01, seed=99).pin_memory()
dev = torch.device("cuda")
adv = adv_h.to(dev).requires_grad_(True); ori = ori_h.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
loss_h = torch.empty((4, B)).pin_memory(); grad_h = torch.empty((B, N, 3)).pin_memory()
def loss_fn(a, o):
    c1, c2 = pcd.distance.chamfer(a, o); h1, h2 = pcd.distance.hausdorff(a, o)
    l = torch.stack([c1, c2, h1, h2]); return l.sum(), (l,)
def timeit(fn, n=20, w=5):
    ts = []
    for k in range(n + w):
        flush.zero_(); torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if k >= w: ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3
g = pcd.graph.GraphedLoss(loss_fn, adv, ori)
def single():
    _, (l,), gr = g.replay(adv_h, ori_h); loss_h.copy_(l, non_blocking=True); grad_h.copy_(gr, non_blocking=True)
print("monolithic e2e: %.1f us" % timeit(single))
l1, g1 = loss_h.clone(), grad_h.clone()
for chunks in (2, 3, 4):
    p = pcd.graph.PipelinedLoss(loss_fn, adv_h, ori_h, chunks=chunks)
    t = timeit(p.replay)
    print("pipelined x%d e2e: %.1f us   same losses %s  grad max rel diff %.1e" % (chunks, t, torch.equal(torch.cat([a[0] for a in p.aux_host], 1), l1), float((p.grad_host - g1).abs().max() / g1.abs().max())))
