"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference (LI-Yiquan/3DPointCloudAttack) is imported from /root/reference with the
work-arounds SURVEY.md section 8c lists (they touch the *environment*, never the
reference's arithmetic):
  * torch.Tensor.cuda -> identity, so the L3 wrappers that hard-code `.cuda()`
    (attack/CW/CW_utils/dist_utils.py:29,68,105,156) run on CPU;
  * stubs for modules missing here (matplotlib, seaborn, open3d) and for the removed
    torch.autograd.gradcheck.zero_gradients, so attack/GeoA3/loss_utils.py imports.

Every vector stores the inputs, the reference outputs and (where the output is
differentiable) the reference autograd gradients for a fixed upstream gradient.
torch 2.11.0+cu128 CPU, fp32, torch.set_num_threads(1) for a fixed summation order.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _prepare_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU box: .cuda() is a no-op
    import importlib
    importlib.import_module("torch.autograd.gradcheck")
    gc = sys.modules["torch.autograd.gradcheck"]
    if not hasattr(gc, "zero_gradients"):
        gc.zero_gradients = lambda x: None
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "open3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["seaborn"].set = lambda *a, **k: None
    import os as _os
    _real_popen = _os.popen

    class _Fake:
        def read(self):
            return "24 80"

    _os.popen = lambda cmd, *a, **k: _Fake() if "stty" in cmd else _real_popen(cmd, *a, **k)


def face_fixture(n, seed):
    """AddData/face0424.txt normalised as pointnet/bosphorus_dataset.py:74-76, random n-subset."""
    raw = np.loadtxt(os.path.join(REF, "AddData", "face0424.txt"), delimiter=",")[:, :3]
    rs = np.random.RandomState(seed)
    pts = raw[rs.permutation(raw.shape[0])[:n]]
    pts = pts - pts.mean(0, keepdims=True)
    pts = pts / np.max(np.sqrt((pts ** 2).sum(1)))
    return pts.astype(np.float32)


def t(a, grad=False):
    x = torch.from_numpy(np.ascontiguousarray(a)).clone()
    x.requires_grad_(grad)
    return x


def n(x):
    return x.detach().cpu().numpy()


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


class _TorchOnCPU:
    """Stands in for the `torch` global of a reference module whose functions hard-code a CUDA
    device (model/dgcnn.py:209 `torch.device('cuda:0')`): every attribute is torch's own, only
    `device(...)` answers cpu.  The reference's arithmetic is untouched."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


def graph_and_sampling():
    """f-2 / f-3: get_graph_feature, LPFA.group_feature, FPS, index_points."""
    rs = np.random.RandomState(20261019)
    ori = np.stack([face_fixture(1024, 0), face_fixture(1024, 1)])
    adv = (ori + 0.01 * rs.randn(*ori.shape)).astype(np.float32)
    adv_cf = np.ascontiguousarray(adv[:, :384].transpose(0, 2, 1))     # graph features on 384 points (fixture size)
    from model import dgcnn, curvenet_util, pointnet2_utils as P2
    out = {}
    saved = dgcnn.torch
    dgcnn.torch = _TorchOnCPU()
    try:
        x_ = t(adv_cf, True)
        f = dgcnn.get_graph_feature(x_, k=20)
        gw = rs.randn(*f.shape).astype(np.float32)
        (f * t(gw)).sum().backward()
        out.update(ggf3=n(f), ggf3_idx=n(dgcnn.knn(t(adv_cf), 20)), ggf3_gw=gw, ggf3_gx=n(x_.grad))
        f16 = rs.randn(2, 16, 128).astype(np.float32)
        x_ = t(f16, True)
        f = dgcnn.get_graph_feature(x_, k=10)
        gw = rs.randn(*f.shape).astype(np.float32)
        (f * t(gw)).sum().backward()
        out.update(f16=f16, ggf16=n(f), ggf16_idx=n(dgcnn.knn(t(f16), 10)), ggf16_gw=gw, ggf16_gx=n(x_.grad))
        # caller-supplied idx with an odd k (scalar path)
        idx7 = rs.randint(0, 128, size=(2, 128, 7)).astype(np.int64)
        out.update(idx7=idx7, ggf16_idx7=n(dgcnn.get_graph_feature(t(f16), k=7, idx=torch.from_numpy(idx7))))
    finally:
        dgcnn.torch = saved
    lp = curvenet_util.LPFA(9, 32, k=20, mlp_num=1, initial=True)
    lp.device = torch.device("cpu")
    xyz_ = t(adv_cf, True)
    pf = lp.group_feature(None if False else t(adv_cf), xyz_, None)
    gw = rs.randn(*pf.shape).astype(np.float32)
    (pf * t(gw)).sum().backward()
    out.update(lpfa9=n(pf), lpfa9_idx=n(curvenet_util.knn(t(adv_cf), 20)[:, :, :20]), lpfa9_gw=gw, lpfa9_gx=n(xyz_.grad))
    # farthest point sampling: random start (pointnet2_utils.py:71) and start 0 (curvenet_util.py:81)
    torch.manual_seed(7)
    st = torch.get_rng_state()
    start = torch.randint(0, 1024, (2,), dtype=torch.long)
    torch.set_rng_state(st)
    out.update(fps_start=n(start), fps_512=n(P2.farthest_point_sample(t(adv), 512)),
               fps0_128=n(curvenet_util.farthest_point_sample(t(adv), 128)))
    out["fps0_all"] = n(curvenet_util.farthest_point_sample(t(adv[:, :300]), 300))   # npoint == N
    fidx = torch.from_numpy(out["fps_512"])
    out["index_points_2d"] = n(P2.index_points(t(adv), fidx))
    bq = P2.query_ball_point(0.2, 32, t(adv), P2.index_points(t(adv), fidx))
    out["ball_idx"] = n(bq)
    out["index_points_3d"] = n(P2.index_points(t(adv), bq[:, :64]))
    # 3-NN inverse-distance interpolation (PointNetFeaturePropagation.forward :289-300), no MLP
    fp = P2.PointNetFeaturePropagation(in_channel=8, mlp=[])
    src = ori[:, ::4][:, :200].copy()                                    # [2,200,3] sources (the unperturbed cloud: no coincident pairs)
    feat = rs.randn(2, 8, 200).astype(np.float32)
    x1, x2, f2 = t(np.ascontiguousarray(adv[:, :600].transpose(0, 2, 1)), True), t(np.ascontiguousarray(src.transpose(0, 2, 1)), True), t(feat, True)
    o = fp(x1, x2, None, f2)                                             # [2,8,600]
    gw = rs.randn(*o.shape).astype(np.float32)
    (o * t(gw)).sum().backward()
    out.update(fp_xyz1=n(x1), fp_xyz2=n(x2), fp_feat=feat, fp_out=n(o), fp_gw=gw, fp_g1=n(x1.grad), fp_g2=n(x2.grad), fp_gf=n(f2.grad))
    save("f_graph_sampling", adv=adv, adv_cf=adv_cf, **out)


def attack_loops():
    """L4: the reference's own CW and kNN attack loops, unmodified, on CPU at B=1
    (their only supported batch size) against tests/tiny_victim.TinyVictim."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    import tiny_victim
    from attack.CW.CW_attack import CW
    from attack.KNN.KNN_attack import CWKNN
    from attack.CW.CW_utils import dist_utils as DU, adv_utils as AU, clip_utils as CU
    victim = tiny_victim.make(seed=3)
    data = face_fixture(256, 5)[None]                                   # [1,256,3]
    with torch.no_grad():
        label = victim(t(data).transpose(1, 2))[0].argmax(1)
    out = dict(tiny_victim.state_to_npz(victim), data=data, label=n(label))
    noises = []
    real_randn = torch.randn

    def recording_randn(*a, **k):
        # The loops start from ori + 1e-7 * randn (CW_attack.py:94): the first Adam steps (update ~ lr * g/|g|)
        # are then decided by gradients at fp32 rounding level, and ANY two correct implementations (the
        # reference on CPU vs. on GPU included) part ways by +-lr per point.  The draws handed to the
        # reference are scaled by 1e4 (start = ori + 1e-3 * randn) so that the comparison is well conditioned;
        # the loop code itself runs unmodified.
        v = real_randn(*a, **k) * 1e4
        noises.append(n(v))
        return v

    tr = lambda f: (lambda a, o, w: f(a.transpose(1, 2).contiguous(), o.transpose(1, 2).contiguous(), w))
    cases = {
        "cw_chamfer": DU.ChamferDist(method="adv2ori"),
        "cw_hausdorff": DU.HausdorffDist(method="ori2adv"),
    }
    torch.randn = recording_randn
    try:
        for tag, df in cases.items():
            noises.clear()
            torch.manual_seed(11)
            atk = CW(victim, victim, AU.UntargetedLogitsAdvLoss(kappa=5.), CU.ClipPointsLinf(budget=0.18), tr(df),
                     attack_lr=1e-2, init_weight=10., max_weight=80., binary_step=3, num_iter=12)
            bestdist, bestattack, _ = atk.attack(t(data), label.clone())
            out.update({tag + "_bestdist": np.asarray(bestdist), tag + "_bestattack": np.asarray(bestattack, np.float32),
                        tag + "_noise": np.stack(noises)})
        # kNN attack (attack/KNN/KNN_attack.py): one long loop, ChamferkNNDist, ProjectInnerClipLinf
        noises.clear()
        torch.manual_seed(12)
        katk = CWKNN(victim, victim, victim, victim, victim, victim, AU.UntargetedLogitsAdvLoss(kappa=15.),
                     DU.ChamferkNNDist(chamfer_method="adv2ori", knn_k=5, knn_alpha=1.05, chamfer_weight=5., knn_weight=3.),
                     CU.ProjectInnerClipLinf(budget=0.1), attack_lr=1e-3, num_iter=20)
        adv, _ = katk.attack(t(data), label.clone())
        out.update(knn_adv=np.asarray(adv, np.float32), knn_noise=np.stack(noises))
    finally:
        torch.randn = real_randn
    save("l4_attack_loops", **out)


def feature_knn():
    """a6 for the feature layers (C = 64, 64, 128): the UNMODIFIED reference DGCNN (model/dgcnn.py:270-311,
    random init, eval mode) runs on two face clouds; every `knn(x, k)` call it makes (dgcnn.py:194-200 via
    get_graph_feature :207) is recorded -- the feature tensor it was given and the index tensor it returned.
    Fixture: N = 512 points, k = 20; the four calls have C = 3, 64, 64, 128 (dgcnn.py:299-311)."""
    import argparse
    from model import dgcnn
    torch.manual_seed(20261020)
    args = argparse.Namespace(k=20, emb_dims=1024, dropout=0.5)
    net = dgcnn.DGCNN(args).eval()
    x = np.stack([face_fixture(512, 3), face_fixture(512, 4)]).transpose(0, 2, 1)
    calls = []
    real_knn = dgcnn.knn

    def recording_knn(x_, k):
        idx = real_knn(x_, k)
        calls.append((n(x_).copy(), n(idx).copy()))
        return idx

    saved = dgcnn.torch
    dgcnn.torch = _TorchOnCPU()
    dgcnn.knn = recording_knn
    try:
        with torch.no_grad():
            net(t(np.ascontiguousarray(x)))
    finally:
        dgcnn.torch = saved
        dgcnn.knn = real_knn
    assert [c[0].shape[1] for c in calls] == [3, 64, 64, 128], [c[0].shape for c in calls]
    out = {}
    for li, (feat, idx) in enumerate(calls):
        out[f"x{li}"] = feat.astype(np.float32)
        out[f"idx{li}"] = idx.astype(np.int16)          # N = 512 fits; halves the fixture
    save("a6_knn_graph_features", **out)


def local_geometry():
    """f-4: attack/GeoA3/utility.py estimate_normal, the brute-force k-NN losses of
    attack/GeoA3/loss_utils.py:107-141 and attack/AOF/TAOF_attack.py get_Laplace_from_pc, run UNMODIFIED.
    Environment work-around (no arithmetic of the reference is touched): `torch.symeig` was removed from torch;
    it is provided as the shim its deprecation note prescribes, torch.linalg.eigh(A, UPLO='U'), and the shim
    records the matrices it is handed so that the Laplacian the reference assembles can be stored."""
    seen = []

    def symeig(A, eigenvectors=False, upper=True):
        seen.append(A.detach().clone())
        e, v = torch.linalg.eigh(A, UPLO="U" if upper else "L")
        return e, v

    torch.symeig = symeig
    from attack.GeoA3 import utility as U
    from attack.GeoA3 import loss_utils as LU
    from attack.AOF import TAOF_attack as TA
    rs = np.random.RandomState(20261021)
    ori = np.stack([face_fixture(512, 5), face_fixture(512, 6)])
    adv = (ori + 0.01 * rs.randn(*ori.shape)).astype(np.float32)
    ori_cf = np.ascontiguousarray(ori.transpose(0, 2, 1)); adv_cf = np.ascontiguousarray(adv.transpose(0, 2, 1))
    out = {"ori": ori_cf, "adv": adv_cf}
    for k in (3, 8, 16):
        seen.clear()
        out[f"normal_k{k}"] = n(U.estimate_normal(t(adv_cf), k))
        out[f"cov_k{k}"] = torch.stack(seen).numpy()                     # [b, n, 3, 3] as the reference formed it
    nrm = out["normal_k8"]
    gN = rs.randn(2, 512).astype(np.float32)
    a_ = t(adv_cf, True)
    v = LU.displacement_loss(a_, t(ori_cf), k=16); (v * t(gN)).sum().backward()
    out["displacement_loss"] = n(v); out["displacement_loss_g"] = n(a_.grad)
    a_ = t(adv_cf, True)
    v = LU.corresponding_normal_loss(a_, t(nrm), k=2); (v * t(gN)).sum().backward()
    out["corresponding_normal_loss"] = n(v); out["corresponding_normal_loss_g"] = n(a_.grad)
    a_ = t(adv_cf, True)
    v = LU.repulsion_loss(a_, k=4, h=0.03); (v * t(gN)).sum().backward()
    out["repulsion_loss"] = n(v); out["repulsion_loss_g"] = n(a_.grad)
    a_ = t(adv_cf, True)
    v = LU.distance_kmean_loss(a_, 8); (v * t(gN)).sum().backward()
    out["distance_kmean_loss"] = n(v); out["distance_kmean_loss_g"] = n(a_.grad)
    # kappa with its gradient (the a4 fixture stores the values; here k = 16 with a gradient w.r.t. the cloud)
    a_ = t(adv_cf, True)
    kap, _ = LU._get_kappa_adv(a_, t(ori_cf), t(nrm), 16); (kap * t(gN)).sum().backward()
    out["kappa_adv_k16"] = n(kap); out["kappa_adv_k16_g"] = n(a_.grad); out["gN"] = gN
    # AOF Laplacian on 256 points (dense [B,N,N]); k = 30 is hard-coded in the reference
    seen.clear()
    small = np.ascontiguousarray(adv_cf[:, :, :256])
    e, vv = TA.get_Laplace_from_pc(t(small))
    out["lap_pc"] = small; out["lap_L"] = seen[0].numpy(); out["lap_e"] = n(e)
    save("f4_local_geometry", **out)


def geoa3_loop():
    """L4: the reference's own GeoA3 attack (attack/GeoA3/GeoA3_attack.py:185-473), unmodified, on CPU
    at B = 1 against tests/tiny_victim.TinyVictim, in two configurations (plain, and projected + clipped offsets).
    Environment work-arounds only: torch.symeig -> torch.linalg.eigh (see local_geometry), the module's bare
    `from utility import ...` needs attack/GeoA3 on sys.path, and nn.init.normal_ is wrapped to RECORD the offset
    draws the loop starts from (the values are passed through unchanged)."""
    import argparse
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    sys.path.insert(0, os.path.join(REF, "attack", "GeoA3"))
    import tiny_victim
    torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO="U" if upper else "L")
    from attack.GeoA3 import GeoA3_attack as GA
    victim = tiny_victim.make(seed=3)
    data = face_fixture(256, 7)[None]                                   # [1,256,3]
    with torch.no_grad():
        label = victim(t(data).transpose(1, 2))[0].argmax(1)
    out = dict(tiny_victim.state_to_npz(victim), data=data, label=n(label))
    draws = []
    real_normal_ = torch.nn.init.normal_

    def recording_normal_(tensor, mean=0., std=1.):
        r = real_normal_(tensor, mean=mean, std=std)
        draws.append(n(r).copy())
        return r

    base = dict(arch="tiny", classes=7, attack_label="Untarget", attack_method="untarget", initial_const=10., lr=0.01,
                optim="adam", binary_max_steps=3, iter_max_steps=15, metric="Loss", cls_loss_type="Margin", confidence=0.,
                dis_loss_type="CD", is_cd_single_side=False, dis_loss_weight=1.0, hd_loss_weight=0.1, curv_loss_weight=1.0,
                curv_loss_knn=16, uniform_loss_weight=0.0, is_pre_jitter_input=False, calculate_project_jitter_noise_iter=50,
                jitter_k=16, jitter_sigma=0.01, jitter_clip=0.05, is_save_normal=False, is_partial_var=False, knn_range=3,
                is_subsample_opt=False, npoint=256, eval_num=1, is_use_lr_scheduler=False, is_pro_grad=False,
                is_real_offset=False, cc_linf=0., is_debug=False, binary_step=3, num_iter=15, output_path="/tmp")
    class _Transposed(torch.nn.Module):
        """the transfer checks at the end of geoA3_attack (:407-471, not on the path) feed [b,n,3] clouds"""

        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            return self.m(x.transpose(2, 1))

    vt = _Transposed(victim)
    GA.nn.init.normal_ = recording_normal_
    try:
        for tag, extra in (("plain", {}), ("proj_clip", dict(is_pro_grad=True, cc_linf=0.02, is_use_lr_scheduler=True))):
            draws.clear()
            torch.manual_seed(21)
            cfg = argparse.Namespace(**dict(base, **extra))
            best_attack, target, ok, best_step, all_loss = GA.geoA3_attack(victim, vt, vt, vt, vt, vt,
                                                                           t(data), label.clone(), cfg, 0, 1)
            out.update({tag + "_best_attack": n(best_attack), tag + "_success": np.asarray(ok), tag + "_best_step": np.asarray(best_step),
                        tag + "_offsets": np.stack(draws), tag + "_loss_n": np.asarray(all_loss, np.float32)})
    finally:
        GA.nn.init.normal_ = real_normal_
    save("l4_geoa3_loop", **out)


def clips():
    """f-1 epilogues (added in round 2): the reference's clip / projection functions, unmodified, on CPU.
    attack/CW/CW_utils/clip_utils.py:5-136 and attack/GeoA3/GeoA3_attack.py:62-101 (module import as in geoa3_loop)."""
    sys.path.insert(0, os.path.join(REF, "attack", "GeoA3"))
    torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO="U" if upper else "L")
    from attack.CW.CW_utils import clip_utils as CU
    from attack.GeoA3 import GeoA3_attack as GA
    rs = np.random.RandomState(77)
    ori = np.ascontiguousarray(np.stack([face_fixture(256, 11), face_fixture(256, 12)]).transpose(0, 2, 1))     # [2,3,256]
    normal = ori - ori.mean(2, keepdims=True) + 0.05 * rs.randn(*ori.shape).astype(np.float32)                  # not unit length
    normal = np.ascontiguousarray(normal.astype(np.float32))
    adv = (ori + 0.03 * rs.randn(*ori.shape)).astype(np.float32)
    adv[0, :, :4] = ori[0, :, :4]                                                    # untouched points
    adv[1, :, 5:9] = ori[1, :, 5:9] - np.float32(0.01) * normal[1, :, 5:9]           # pushed straight inside: the "opposite" branch
    adv[1, :, 9:12] = ori[1, :, 9:12] + np.float32(1e-8)                             # shorter than every epsilon
    out = dict(ori=ori, adv=adv, normal=normal)
    out["linf"] = n(CU.ClipPointsLinf(budget=0.03)(t(adv), t(ori)))
    out["l2"] = n(CU.ClipPointsL2(budget=0.5)(t(adv), t(ori)))
    out["project_linf"] = n(CU.ProjectInnerClipLinf(budget=0.03)(t(adv), t(ori), t(normal)))
    off = adv - ori
    out["lp_clip"] = n(GA.lp_clip(t(off), 0.02))
    out["offset_proj"] = n(GA.offset_proj(t(off), t(ori), t(normal)))
    out["find_offset"] = n(GA.find_offset(t(ori), t(adv)))
    save("f1_clips", **out)


def main():
    torch.set_num_threads(1)
    _prepare_reference()
    if "--clips" in sys.argv:          # f-1 clip / projection epilogues (added in round 2)
        clips()
        return
    if "--geoa3" in sys.argv:          # the reference's GeoA3 loop at B = 1 (added in round 2)
        geoa3_loop()
        return
    if "--geometry" in sys.argv:       # f-4 (added in round 2)
        local_geometry()
        return
    if "--features" in sys.argv:       # a6 for C = 64 / 128 (added in round 2)
        feature_knn()
        return
    if "--loops" in sys.argv:
        attack_loops()
        return
    if "--graph" in sys.argv:          # only the section added after the first fixtures were committed
        graph_and_sampling()
        return
    rs = np.random.RandomState(20261018)

    face = face_fixture(1024, 0)                                  # [1024,3]
    ori = np.stack([face, face_fixture(1024, 1)])                 # [2,1024,3]
    adv = (ori + 0.01 * rs.randn(*ori.shape)).astype(np.float32)
    adv0 = (ori + 1e-7 * rs.randn(*ori.shape)).astype(np.float32)  # CW iteration-0 state
    gB = np.array([0.7, -1.3], np.float32)
    gB2 = np.array([1.9, 0.4], np.float32)

    # ---------------- a1: utils/dis_utils_torch.py ------------------------------------
    from utils import dis_utils_torch as D
    a_cf = np.ascontiguousarray(adv.transpose(0, 2, 1)); b_cf = np.ascontiguousarray(ori.transpose(0, 2, 1))
    out = {}
    for fn in ("chamfer", "sgd_hausdorff_dis", "bid_hausdorff_dis"):
        a_, b_ = t(a_cf, True), t(b_cf, True)
        v = getattr(D, fn)(a_, b_)
        v.backward()
        out[fn] = n(v); out[fn + "_ga"] = n(a_.grad); out[fn + "_gb"] = n(b_.grad)
    out["pairwise_distances"] = n(D.pairwise_distances(t(a_cf[:, :, :96]), t(b_cf[:, :, :80])))
    # the commented-out 3-point example at utils/dis_utils_torch.py:30-35
    ka = np.array([[[1, 1, 1], [1, 1, 1], [1, 1, 1]]], np.float32)
    kb = np.array([[[2, 2, 3], [2, 2, 2], [2, 2, 2]]], np.float32)
    out["ka"] = ka; out["kb"] = kb
    out["k_euclidean"] = n(D.euclidean_distances(t(ka), t(kb)))
    out["k_chamfer"] = n(D.chamfer(t(ka), t(kb)))
    out["k_sgd"] = n(D.sgd_hausdorff_dis(t(ka), t(kb)))
    out["k_bid"] = n(D.bid_hausdorff_dis(t(ka), t(kb)))
    save("a1_dis_utils_torch", a=a_cf, b=b_cf, **out)

    # ---------------- a2: attack/CW/CW_utils/distance.py -------------------------------
    from attack.CW.CW_utils import distance as CD
    for tag, P_, G_ in (("face", adv, ori), ("iter0", adv0, ori),
                        ("ragged", adv[:, :700], ori[:, :1000])):
        out = {}
        for name, mod in (("chamfer", CD.chamfer), ("hausdorff", CD.hausdorff)):
            p_, g_ = t(P_, True), t(G_, True)
            l1, l2 = mod(p_, g_)
            ((l1 * t(gB)).sum() + (l2 * t(gB2)).sum()).backward()
            out[name + "_l1"] = n(l1); out[name + "_l2"] = n(l2)
            out[name + "_gp"] = n(p_.grad); out[name + "_gg"] = n(g_.grad)
        Pm = CD.chamfer.batch_pairwise_dist(t(G_), t(P_))            # [B, N2(gts), N1(preds)]
        m1, i1 = torch.min(Pm, 1); m2, i2 = torch.min(Pm, 2)
        out.update(col_min=n(m1), col_arg=n(i1), row_min=n(m2), row_arg=n(i2),
                   P_block=n(Pm[:, :64, :48]))
        save("a2_distance_" + tag, preds=P_, gts=G_, g1=gB, g2=gB2, **out)

    # SIadv variant returns (loss1+loss2)/2
    from attack.SIadv.utils import set_distance as SD
    save("a2_set_distance", preds=adv, gts=ori,
         chamfer=n(SD.chamfer(t(adv), t(ori))), hausdorff=n(SD.hausdorff(t(adv), t(ori))))

    # tie case: duplicated points -> torch.min(dim) picks the first index
    dup_g = ori[:, :256].copy(); dup_g[:, 128:] = dup_g[:, :128]
    dup_p = adv[:, :256].copy(); dup_p[:, 64:128] = dup_p[:, :64]
    Pm = CD.chamfer.batch_pairwise_dist(t(dup_g), t(dup_p))
    m1, i1 = torch.min(Pm, 1); m2, i2 = torch.min(Pm, 2)
    save("a2_distance_ties", preds=dup_p, gts=dup_g, col_min=n(m1), col_arg=n(i1), row_min=n(m2), row_arg=n(i2))

    # ---------------- L3 wrappers: attack/CW/CW_utils/dist_utils.py --------------------
    from attack.CW.CW_utils import dist_utils as DU
    w = np.array([2.0, 0.5], np.float32)
    out = {}
    for method in ("adv2ori", "ori2adv", "avg"):
        for cls in ("ChamferDist", "HausdorffDist"):
            a_ = t(adv, True)
            v = getattr(DU, cls)(method=method)(a_, t(ori), weights=t(w), batch_avg=False)
            (v * t(gB)).sum().backward()
            out[f"{cls}_{method}"] = n(v); out[f"{cls}_{method}_g"] = n(a_.grad)
    a_ = t(adv, True)
    v = DU.ChamferDist()(a_, t(ori)); v.backward()
    out["ChamferDist_default_mean"] = n(v); out["ChamferDist_default_mean_g"] = n(a_.grad)
    for k, alpha in ((5, 1.05), (16, 1.05)):
        a_ = t(adv, True)
        v = DU.KNNDist(k=k, alpha=alpha)(a_, weights=t(w), batch_avg=False)
        (v * t(gB)).sum().backward()
        out[f"KNNDist_k{k}"] = n(v); out[f"KNNDist_k{k}_g"] = n(a_.grad)
    a_ = t(adv, True)
    v = DU.ChamferkNNDist(knn_k=16)(a_, t(ori), weights=t(w), batch_avg=True); v.backward()
    out["ChamferkNNDist_k16"] = n(v); out["ChamferkNNDist_k16_g"] = n(a_.grad)
    save("l3_dist_utils", adv=adv, ori=ori, weights=w, gB=gB, **out)

    # ---------------- a3: attack/GeoA3/knn_utils.py -------------------------------------
    from attack.GeoA3 import knn_utils as KU
    out = {}
    for tag, p1, p2, K in (("cross1", adv, ori, 1), ("self17", adv, adv, 17), ("cross4", ori, adv, 4)):
        a_, b_ = t(p1, True), (None if p2 is p1 else t(p2, True))
        r = KU.knn_points(a_, a_ if b_ is None else b_, K=K, return_nn=True)
        gw = t(rs.randn(*r.dists.shape).astype(np.float32))
        (r.dists * gw).sum().backward()
        out.update({tag + "_dists": n(r.dists), tag + "_idx": n(r.idx), tag + "_nn": n(r.knn),
                    tag + "_gw": n(gw), tag + "_g1": n(a_.grad)})
        if b_ is not None:
            out[tag + "_g2"] = n(b_.grad)
    feat = rs.randn(2, 1024, 5).astype(np.float32)
    idx = out["self17_idx"]
    out["gather_x"] = feat
    out["gather_out"] = n(KU.knn_gather(t(feat), torch.from_numpy(idx)))
    save("a3_knn_utils", adv=adv, ori=ori, **out)

    # ---------------- a4: attack/GeoA3/loss_utils.py ------------------------------------
    from attack.GeoA3 import loss_utils as LU
    adv_cf = np.ascontiguousarray(adv.transpose(0, 2, 1)); ori_cf = np.ascontiguousarray(ori.transpose(0, 2, 1))
    nrm = rs.randn(2, 3, 1024).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    out = {}
    for fn in ("chamfer_loss", "pseudo_chamfer_loss", "hausdorff_loss"):
        a_ = t(adv_cf, True)
        v = getattr(LU, fn)(a_, t(ori_cf)); (v * t(gB)).sum().backward()
        out[fn] = n(v); out[fn + "_g"] = n(a_.grad)
    a_ = t(adv_cf, True)
    v = LU.kNN_smoothing_loss(a_, 16); (v * t(gB)).sum().backward()
    out["kNN_smoothing_loss"] = n(v); out["kNN_smoothing_loss_g"] = n(a_.grad)
    ori_kappa = LU._get_kappa_ori(t(ori_cf), t(nrm), 16)
    a_ = t(adv_cf, True)
    adv_kappa, normal_curr = LU._get_kappa_adv(a_, t(ori_cf), t(nrm), 16)
    v = LU.curvature_loss(a_, t(ori_cf), adv_kappa, ori_kappa); (v * t(gB)).sum().backward()
    out.update(ori_kappa=n(ori_kappa), adv_kappa=n(adv_kappa), normal_curr=n(normal_curr),
               curvature_loss=n(v), curvature_loss_g=n(a_.grad))
    save("a4_geoa3_losses", adv=adv_cf, ori=ori_cf, normal=nrm, gB=gB, **out)

    # ---------------- a6: model/dgcnn.py, model/curvenet_util.py ------------------------
    from model import dgcnn, curvenet_util
    f64c = rs.randn(2, 64, 256).astype(np.float32)
    save("a6_knn_graph", x3=adv_cf, f64=f64c,
         dgcnn_k20=n(dgcnn.knn(t(adv_cf), 20)), dgcnn_f64_k20=n(dgcnn.knn(t(f64c), 20)),
         curvenet_k20=n(curvenet_util.knn(t(adv_cf), 20)),
         curvenet_normal_k20=n(curvenet_util.normal_knn(t(adv_cf), 20)))

    # ---------------- a7: model/pointnet2_utils.py --------------------------------------
    from model import pointnet2_utils as P2
    xyz = adv; new_xyz = adv[:, ::2][:, :512].copy()
    save("a7_pointnet2_utils", xyz=xyz, new_xyz=new_xyz,
         sqdist_block=n(P2.square_distance(t(new_xyz[:, :64]), t(xyz[:, :96]))),
         ball_r02_n32=n(P2.query_ball_point(0.2, 32, t(xyz), t(new_xyz))),
         ball_r04_n64=n(P2.query_ball_point(0.4, 64, t(xyz[:, :512]), t(new_xyz[:, :128]))),
         ball_r002_n8=n(P2.query_ball_point(0.02, 8, t(xyz), t(new_xyz[:, :64]))))


if __name__ == "__main__":
    main()
