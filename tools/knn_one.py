"""one k-NN call (profiling target): python tools/knn_one.py B N C K"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pcd = importlib.import_module("3dpointcloudattack_b200")
B, N, C, K = [int(a) for a in sys.argv[1:5]]
x = torch.rand(B, C, N, device="cuda") if C == 3 else torch.randn(B, C, N, device="cuda")
pts = x.transpose(1, 2)
for _ in range(2):
    d, i = pcd.functional.knn(pts, pts, K)
torch.cuda.synchronize()
print("ok", float(d.sum()))
