"""Where does a CW iteration spend its time (development tool)?  eager vs CUDA graph, per kernel."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import victims
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda")
B, N = 32, 4096
torch.manual_seed(0)
model = victims.PointNetVictim(106).to(dev).eval()
ori = synth.face_clouds(B, N, seed=1234).to(dev)
target = torch.arange(B, device=dev) % 106
cd, hd = pcd.dist_utils.ChamferDist(method="avg"), pcd.dist_utils.HausdorffDist(method="avg")
dist = lambda a, o, w, batch_avg=False: cd(a, o, weights=w, batch_avg=batch_avg) + hd(a, o, weights=w, batch_avg=batch_avg)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for mode in ("eager", "graph"):
    atk = pcd.cw_loop.CWAttack(model, pcd.cw_loop.UntargetedLogitsAdvLoss(kappa=30.), dist, num_iter=20, binary_step=1,
                               clip_func=pcd.cw_loop.ClipPointsLinf(0.18), use_graph=(mode == "graph"))
    atk.attack(ori, target, seed=1)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        atk.attack(ori, target, seed=2)
    print("=====", mode, "loop ms/iter", atk.loop_ms / 20)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))

# victim alone
x = ori.transpose(1, 2).contiguous().requires_grad_(True)
def fb():
    x.grad = None
    model(x)[0].sum().backward()
print("victim fwd+bwd eager ms", timeit(fb))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): fb()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
with torch.cuda.graph(g, stream=s):
    fb()
print("victim fwd+bwd graph ms", timeit(g.replay))
