"""Flagged (near-tie) points of the approximate sweep per side (development tool; needs a -DPCD_APX_DEBUG build)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "tools", "libpcdist_apxdbg.so")
os.environ["PCDIST_LIBRARY"] = out                      # before the package is imported (it reads the variable at import)
import importlib.util
spec = importlib.util.spec_from_file_location("_pcd_build", os.path.join(ROOT, "3dpointcloudattack_b200", "build.py"))
build = importlib.util.module_from_spec(spec); spec.loader.exec_module(build)
if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(ROOT, "3dpointcloudattack_b200", "csrc", "pcd_nn1.cu")):
    build.build(force=True, extra_flags=["-DPCD_APX_DEBUG"], out=out)
import torch
pcd = importlib.import_module("3dpointcloudattack_b200"); F = pcd.functional
synth = importlib.import_module("3dpointcloudattack_b200.synth")
F._keep_workspace = True
def al(v, a): return (v + a - 1) // a * a
for (B, N, sigma) in [(32, 4096, 0.01), (32, 4096, 1e-7), (8, 4096, 0.01), (32, 1024, 0.01), (32, 4096, 1.0)]:
    ori = synth.face_clouds(B, N, seed=1234).cuda()
    adv = (ori + sigma * torch.randn_like(ori)).contiguous()
    F.force_sweep_mode(F.SWEEP_APPROX)
    r = F.nn1(adv, ori, F.FORM_SUM_FIRST, F.NORM_FMA, cache=False)
    torch.cuda.synchronize()
    Npad, Mpad = al(N, 2048), al(N, 256)
    off = al(B * Npad * 8 + B * Mpad * 8 + B * 8, 16)
    cnt = F._last_workspace[off + 4 * (B + 1): off + 4 * (B + 1) + 8].view(torch.int32).tolist()
    norms = (ori ** 2).sum(-1)
    print(f"B={B} N={N} sigma={sigma}: flagged rows {cnt[0]} / {B * N}  columns {cnt[1]} / {B * N}   max |p|^2 {float(norms.max()):.3f}  "
          f"median row min {float(r.row_min.median()):.3e}", flush=True)
