"""One GeoA3 geometry-loss forward+backward + normals + k-NN outlier loss at configs[3] shard size (profiling target for ncu)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
LU = pcd.loss_utils
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ori = synth.face_clouds(min(B, 8), N, seed=3).cuda().repeat((B + 7) // 8, 1, 1)[:B].transpose(1, 2).contiguous()
for _ in range(2):
    normal = pcd.utility.estimate_normal(ori, 3)
    kappa_ori = LU._get_kappa_ori(ori, normal, 16)
    adv = (ori + 0.01 * torch.randn_like(ori)).requires_grad_(True)
    cd = LU.chamfer_loss(adv, ori); hd = LU.hausdorff_loss(adv, ori)
    kap, _ = LU._get_kappa_adv(adv, ori, normal, 16)
    (cd + 0.1 * hd + LU.curvature_loss(adv, ori, kap, kappa_ori) + LU.kNN_smoothing_loss(adv, 16)).sum().backward()
torch.cuda.synchronize()
print("ok")
