"""Host-to-host step time of PipelinedLoss for slice layouts / priorities (development tool)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
ori_h = synth.face_clouds(32, 4096, seed=1234).pin_memory(); adv_h = synth.perturb(ori_h, 0.01, seed=99).pin_memory()
def loss_fn(a, o):
    c1, c2 = pcd.distance.chamfer(a, o); h1, h2 = pcd.distance.hausdorff(a, o)
    l = torch.stack([c1, c2, h1, h2]); return l.sum(), (l,)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def measure(**kw):
    piped = pcd.graph.PipelinedLoss(loss_fn, adv_h, ori_h, **kw)
    for _ in range(5): piped.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); piped.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]
print("priority range", torch.cuda.Stream.priority_range())
cfgs = [[8, 8, 8, 8], [6, 10, 10, 6], [8, 12, 12], [7, 12, 13], [6, 13, 13], [8, 11, 13], [10, 11, 11], [8, 10, 14], [6, 8, 9, 9]]
objs = [pcd.graph.PipelinedLoss(loss_fn, adv_h, ori_h, slice_sizes=c) for c in cfgs]
for o in objs:
    for _ in range(5): o.replay()
torch.cuda.synchronize()
res = {i: [] for i in range(len(cfgs))}
for rnd in range(40):
    for i, o in enumerate(objs):
        flush.zero_(); torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o.replay(); e1.record(); torch.cuda.synchronize()
        res[i].append(e0.elapsed_time(e1) * 1e3)
for i, c in enumerate(cfgs):
    v = sorted(res[i])
    print(f"{c!s:22s} median {v[len(v)//2]:7.1f} us  best {v[0]:7.1f}  p90 {v[int(len(v)*0.9)]:7.1f}", flush=True)
