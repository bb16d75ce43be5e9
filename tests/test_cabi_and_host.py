"""CPU-only checks: the C-ABI library builds, loads and exports exactly what include/pcdist.h
declares; the product path fails loudly without CUDA; host-side logic (installer, sharding)."""
import importlib
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pcd = importlib.import_module("3dpointcloudattack_b200")


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pcdist.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = pcd._lib.load()
    syms = header_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(pcd._lib.SIGNATURES) == syms           # the ctypes table covers the header, no more no less
    out = subprocess.check_output(["nm", "-D", "--defined-only", pcd._lib.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r" T (pcd_[a-z0-9_]+)", out)))
    assert exported == syms
    assert lib.pcd_version() == 200


def test_library_is_sm100a_and_uses_blackwell_instructions():
    out = subprocess.run(["cuobjdump", "-sass", pcd._lib.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out
    for mnemonic in ("FFMA2", "FADD2", "FMUL2", "FMNMX3", "CREDUX", "UBLKCP"):
        assert mnemonic in out, mnemonic


def test_workspace_sizes_without_gpu():
    lib = pcd._lib.load()
    assert lib.pcd_nn1_workspace_bytes(32, 4096, 4096) >= 32 * 4096 * (16 + 16 + 8 + 8)
    assert lib.pcd_nn1_workspace_bytes(0, 1, 1) == 0
    assert lib.pcd_knn_workspace_bytes(2, 100, 100, 3, 5) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_no_cpu_fallback():
    a = torch.zeros(1, 8, 3)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.distance.chamfer(a, a)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.knn_utils.knn_points(a, a, K=2)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.pointnet2_utils.query_ball_point(0.2, 4, a, a)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.dgcnn.knn(a.transpose(1, 2), 2)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.dgcnn.get_graph_feature(a.transpose(1, 2).contiguous(), k=2, idx=torch.zeros(1, 8, 2, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.pointnet2_utils.farthest_point_sample(a, 4)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.curvenet_util.farthest_point_sample(a, 4)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pcd.cw_loop.CWAttack(torch.nn.Identity(), None, None).attack(a, torch.zeros(1))
    with pytest.raises(ValueError, match="pinned"):
        pcd.graph.PipelinedLoss(lambda x, y: (x.sum(), ()), a, a)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "3dpointcloudattack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", ""), f


def test_signatures_mirror_the_reference():
    import inspect
    assert list(inspect.signature(pcd.knn_utils.knn_points).parameters) == [
        "p1", "p2", "lengths1", "lengths2", "K", "version", "return_nn", "return_sorted"]
    assert list(inspect.signature(pcd.knn_utils.knn_gather).parameters) == ["x", "idx", "lengths"]
    assert list(inspect.signature(pcd.pointnet2_utils.query_ball_point).parameters) == ["radius", "nsample", "xyz", "new_xyz"]
    assert list(inspect.signature(pcd.dist_utils.ChamferkNNDist.__init__).parameters) == [
        "self", "chamfer_method", "knn_k", "knn_alpha", "chamfer_weight", "knn_weight"]
    assert list(inspect.signature(pcd.dist_utils.KNNDist.forward).parameters) == ["self", "pc", "weights", "batch_avg"]
    assert list(inspect.signature(pcd.dist_utils.ChamferDist.forward).parameters) == ["self", "adv_pc", "ori_pc", "weights", "batch_avg"]
    assert list(inspect.signature(pcd.dgcnn.get_graph_feature).parameters) == ["x", "k", "idx"]
    for fn in ("euclidean_distances", "pairwise_distances", "chamfer", "sgd_hausdorff_dis", "bid_hausdorff_dis"):
        assert callable(getattr(pcd.dis_utils_torch, fn))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_install_patches_reference_modules():
    sys.path.insert(0, "/root/reference")
    try:
        rep = pcd.install.install(modules=["utils.dis_utils_torch", "attack.CW.CW_utils.distance",
                                           "attack.CW.CW_utils.dist_utils", "attack.GeoA3.knn_utils",
                                           "model.dgcnn", "model.pointnet2_utils"])
        assert all(v == "patched" for v in rep.values()), rep
        import attack.CW.CW_utils.dist_utils as DU
        import attack.CW.CW_utils.distance as CD
        import model.dgcnn as MD
        assert CD.chamfer is pcd.distance.chamfer and DU.chamfer is pcd.distance.chamfer
        assert DU.ChamferDist is pcd.dist_utils.ChamferDist and MD.knn is pcd.dgcnn.knn
        assert hasattr(DU, "L2Dist")                       # untouched names stay the reference's
    finally:
        sys.path.remove("/root/reference")
        for m in [m for m in sys.modules if m.split(".")[0] in ("attack", "model", "utils")]:
            del sys.modules[m]
