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
for tag, kw in (("plain", {}), ("proj_clip", dict(is_pro_grad=True, cc_linf=0.02, is_use_lr_scheduler=True))):
    atk = G.GeoA3Attack(victim, classes=7, initial_const=10., lr=0.01, binary_max_steps=3, iter_max_steps=15, **kw)
    rec = []
    orig = atk._iteration
    def it(st, orig=orig, rec=rec):
        orig(st); rec.append(float(st["loss_n"][0]))
    atk._iteration = it
    data = torch.from_numpy(g["data"]).cuda(); label = torch.from_numpy(g["label"]).cuda()
    best, ok, bl, bs = atk.attack(data, label, init_offset=torch.from_numpy(g[tag + "_offsets"]))
    ours = np.array(rec[-15:]); ref = g[tag + "_loss_n"][:, 0]
    print(tag, "loss_n last search step: ours vs ref")
    for a, b in zip(ours, ref): print(f"   {a: .7f} {b: .7f}  diff {a-b: .2e}")
    ours0 = np.array(rec[:15]); print("  first search step ours:", ours0[:4])
