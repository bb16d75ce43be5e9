import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pcd = importlib.import_module("3dpointcloudattack_b200")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
import tiny_victim
g = np.load(os.path.join(ROOT, "tests/golden/l4_geoa3_loop.npz"))
victim = tiny_victim.from_npz(g).cuda()
G = pcd.geoa3_loop
data = torch.from_numpy(g["data"]).cuda(); label = torch.from_numpy(g["label"]).cuda()
res = {}
for name, kw in (("eager", {}), ("graph", dict(use_graph=True))):
    for bs in (1, 3):
        atk = G.GeoA3Attack(victim, classes=7, initial_const=10., lr=0.01, binary_max_steps=bs, iter_max_steps=15, **kw)
        best, ok, bl, bstep = atk.attack(data, label, init_offset=torch.from_numpy(g["plain_offsets"]))
        res[(name, bs)] = (best.cpu().numpy(), bl.cpu().numpy(), bstep.cpu().numpy())
        print(name, bs, "best_loss", bl.cpu().numpy(), "best_step", bstep.cpu().numpy())
for bs in (1, 3):
    d = np.abs(res[("eager", bs)][0] - res[("graph", bs)][0])
    print("binary steps", bs, "eager vs graph: max", d.max(), "median", np.median(d), "frac<1e-5", (d < 1e-5).mean())
