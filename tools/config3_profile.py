"""Per-kernel device time of one kNN-attack iteration vs the PointNet++ SSG victim (development tool)."""
import importlib, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
import victims
CL = pcd.cw_loop
torch.manual_seed(0)
v = victims.PointNet2SSGVictim(pcd.pointnet2_utils.sample_and_group).cuda().eval()
data = synth.face_clouds(64, 1024, seed=77).cuda()
with torch.no_grad():
    target = v(data.transpose(1, 2))[0].argmax(1)
atk = CL.KNNAttack(v, CL.UntargetedLogitsAdvLoss(kappa=15.), pcd.dist_utils.ChamferkNNDist(knn_k=16), CL.ProjectInnerClipLinf(0.1), attack_lr=1e-3, num_iter=5)
atk.attack(data, target)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    atk.attack(data, target)
    torch.cuda.synchronize()
tot = sum(e.device_time_total for e in prof.key_averages())
print("total device time per iteration %.2f ms" % (tot / 5 / 1e3))
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:16]:
    print(f"  {e.device_time_total / 5:9.1f} us/iter  x{e.count / 5:5.1f}  {e.key[:100]}")
