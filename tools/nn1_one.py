"""A few eager Chamfer+Hausdorff forward+backward steps at BASELINE configs[1] (profiling target for ncu)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ori = synth.face_clouds(min(B, 8), N, seed=1234).cuda().repeat((B + 7) // 8, 1, 1)[:B].contiguous()
adv = (ori + 0.01 * torch.randn_like(ori)).requires_grad_(True)
for _ in range(reps):
    adv.grad = None
    c1, c2 = pcd.distance.chamfer(adv, ori); h1, h2 = pcd.distance.hausdorff(adv, ori)
    torch.stack([c1, c2, h1, h2]).sum().backward()
torch.cuda.synchronize()
print("ok", float(c1.sum()))
