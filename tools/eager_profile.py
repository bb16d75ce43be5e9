"""cProfile of the eager Chamfer+Hausdorff forward+backward (host side), development tool."""
import cProfile, importlib, os, pstats, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
ori = synth.face_clouds(4, 1024, seed=1).cuda(); adv = synth.perturb(ori.cpu(), 0.01, seed=2).cuda().requires_grad_(True)
def fb():
    adv.grad = None
    c1, c2 = pcd.distance.chamfer(adv, ori); h1, h2 = pcd.distance.hausdorff(adv, ori)
    (c1 + c2 + h1 + h2).sum().backward()
for _ in range(50): fb()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(1000): fb()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
