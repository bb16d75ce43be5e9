"""Host-to-host step (PipelinedLoss) for several chunk counts (development tool)."""
import importlib, os, sys
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
for chunks in (1, 2, 4, 8):
    piped = pcd.graph.PipelinedLoss(loss_fn, adv_h, ori_h, chunks=chunks)
    ms = []
    for k in range(25):
        flush.zero_(); torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); piped.replay(); e1.record(); torch.cuda.synchronize()
        if k >= 5: ms.append(e0.elapsed_time(e1))
    ms.sort()
    print(f"chunks={chunks}: median {ms[len(ms)//2]*1e3:.1f} us  min {ms[0]*1e3:.1f}")
