"""Host-side cost of the eager path (development tool): CPU time per call with the GPU kept far behind."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pcd = importlib.import_module("3dpointcloudattack_b200")
synth = importlib.import_module("3dpointcloudattack_b200.synth")
F = pcd.functional
ori = synth.face_clouds(4, 1024, seed=1).cuda(); adv = synth.perturb(ori.cpu(), 0.01, seed=2).cuda().requires_grad_(True)
def cpu_time(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return t
print("nn1 forward (no grad)          %.1f us" % cpu_time(lambda: F.nn1(ori, adv.detach(), F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)))
print("nn1 forward (grad)             %.1f us" % cpu_time(lambda: F.nn1(ori, adv, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)))
def fb():
    adv.grad = None
    c1, c2 = pcd.distance.chamfer(adv, ori); h1, h2 = pcd.distance.hausdorff(adv, ori)
    (c1 + c2 + h1 + h2).sum().backward()
print("chamfer+hausdorff fwd+bwd      %.1f us" % cpu_time(fb))
def fonly():
    c1, c2 = pcd.distance.chamfer(adv, ori); h1, h2 = pcd.distance.hausdorff(adv, ori)
    return (c1 + c2 + h1 + h2).sum()
print("chamfer+hausdorff fwd only     %.1f us" % cpu_time(fonly))
print("torch.empty x7                 %.1f us" % cpu_time(lambda: [torch.empty((4, 1024), device="cuda") for _ in range(7)]))
lib = pcd._lib.load()
print("ctypes pcd_version call        %.2f us" % cpu_time(lambda: lib.pcd_version(), n=2000))
x = torch.zeros(8, device="cuda", requires_grad=True)
print("trivial torch fwd+bwd          %.1f us" % cpu_time(lambda: (x * 2).sum().backward()))
