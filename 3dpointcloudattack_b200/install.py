"""Install the B200 kernels behind the reference's own import paths.

    import sys; sys.path.insert(0, "/path/to/3DPointCloudAttack")
    from pcdist import install; install.install()

The reference binds names at import time (`from attack.CW.CW_utils.distance import chamfer,
hausdorff` in dist_utils.py:6; `from attack.GeoA3.knn_utils import knn_points, knn_gather` in
loss_utils.py:14), so both the defining modules and the already-imported dependants are
patched.  Modules that cannot be imported (missing open3d etc.) are skipped and reported.
"""
import importlib

from . import (curvenet_util, cw_loop, dgcnn, dis_utils_torch, dist_utils, distance, geoa3_loop, knn_utils, loss_utils,
               pointnet2_utils, set_distance, taof, utility)

# reference module -> (our module, names)
PATCHES = {
    "utils.dis_utils_torch": (dis_utils_torch, ["chamfer", "sgd_hausdorff_dis", "bid_hausdorff_dis"]),
    "attack.CTA.utils.dis_utils_torch": (dis_utils_torch, ["chamfer", "sgd_hausdorff_dis", "bid_hausdorff_dis"]),
    "attack.CW.CW_utils.distance": (distance, ["ChamferDistance", "HausdorffDistance", "chamfer", "hausdorff"]),
    "attack.Gen3DAdv.utils.distance": (distance, ["ChamferDistance", "HausdorffDistance", "chamfer", "hausdorff"]),
    "attack.SIadv.utils.set_distance": (set_distance, ["ChamferDistance", "HausdorffDistance", "chamfer", "hausdorff"]),
    "attack.CW.CW_utils.dist_utils": (dist_utils, ["ChamferDist", "HausdorffDist", "KNNDist", "ChamferkNNDist"]),
    "attack.Gen3DAdv.utils.dist_utils": (dist_utils, ["ChamferDist", "HausdorffDist", "KNNDist", "ChamferkNNDist"]),
    "attack.SIadv.baselines.attack.util.dist_utils": (dist_utils, ["ChamferDist", "HausdorffDist", "KNNDist", "ChamferkNNDist"]),
    "attack.CW.CW_utils.clip_utils": (cw_loop, ["ClipPointsL2", "ClipPointsLinf", "ProjectInnerClipLinf"]),   # one launch each
    "attack.GeoA3.knn_utils": (knn_utils, ["knn_points", "knn_gather"]),
    "attack.GeoA3.loss_utils": (loss_utils, ["knn_points", "knn_gather", "chamfer_loss", "pseudo_chamfer_loss",
                                            "hausdorff_loss", "_get_kappa_ori", "_get_kappa_adv", "curvature_loss",
                                            "kNN_smoothing_loss", "displacement_loss", "corresponding_normal_loss",
                                            "repulsion_loss", "distance_kmean_loss"]),
    "attack.GeoA3.GeoA3_attack": (knn_utils, ["knn_points", "knn_gather"]),
    "attack.GeoA3.utility": (utility, ["estimate_normal", "estimate_perpendicular", "estimate_normal_via_ori_normal"]),
    "model.dgcnn": (dgcnn, ["knn", "get_graph_feature"]),
    "pointnet.model": (dgcnn, ["knn", "get_graph_feature"]),
    "attack.AOF.TAOF_attack": (taof, ["knn", "get_Laplace_from_pc"]),   # knn = dgcnn.knn's formulation (TAOF_attack.py:13-28)
    "attack.AOF.Eval_AOF": (taof, ["knn", "get_Laplace_from_pc"]),
    "model.curvenet_util": (curvenet_util, ["knn", "normal_knn", "farthest_point_sample"]),
    "model.pointnet2_utils": (pointnet2_utils, ["query_ball_point", "farthest_point_sample"]),
    "pointnet.pointnet2_utils": (pointnet2_utils, ["query_ball_point", "farthest_point_sample"]),
}


def install(modules=None, strict=False):
    """Patch the reference modules listed in PATCHES (or the subset `modules`).
    Returns {module_name: "patched" | "skipped: <import error>"}."""
    report = {}
    for name, (ours, names) in PATCHES.items():
        if modules is not None and name not in modules:
            continue
        try:
            ref = importlib.import_module(name)
        except Exception as e:                     # missing open3d / removed torch APIs ...
            if strict:
                raise
            report[name] = f"skipped: {type(e).__name__}: {e}"
            continue
        for n in names:
            if hasattr(ours, n):
                setattr(ref, n, getattr(ours, n))
        if name in ("model.pointnet2_utils", "pointnet.pointnet2_utils") and hasattr(ref, "PointNetFeaturePropagation"):
            ref.PointNetFeaturePropagation.forward = pointnet2_utils.feature_propagation_forward    # 3-NN interpolation :273-311
        if name == "model.curvenet_util":
            ref.query_ball_point = pointnet2_utils.query_ball_point          # copy at curvenet_util.py:93-113
            ref.LPFA.group_feature = curvenet_util.group_feature             # method :206-236
        if name in ("attack.GeoA3.utility", "attack.GeoA3.GeoA3_attack"):
            ref.knn_points, ref.knn_gather = knn_utils.knn_points, knn_utils.knn_gather
        if name == "attack.GeoA3.GeoA3_attack":
            for n in ("estimate_normal", "estimate_perpendicular", "estimate_normal_via_ori_normal"):
                if hasattr(ref, n):
                    setattr(ref, n, getattr(utility, n))
            for n in ("offset_proj", "find_offset", "lp_clip"):              # GeoA3_attack.py:62-101, fused epilogues
                setattr(ref, n, getattr(geoa3_loop, n))
        if name.endswith("dist_utils") and hasattr(ref, "chamfer"):
            ref.chamfer, ref.hausdorff = distance.chamfer, distance.hausdorff
        report[name] = "patched"
    return report
