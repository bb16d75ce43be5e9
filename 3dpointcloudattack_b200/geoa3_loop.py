"""Device-resident, batch-safe GeoA3 attack loop (the caller of the path).

Same optimisation as attack/GeoA3/GeoA3_attack.py:185-404 in its standard configuration (whole-cloud offset
variable, no input jitter, no sub-sampling): Adam / SGD on a per-point offset, loss = classification margin
+ scale_const * (w_cd * Chamfer + w_hd * Hausdorff + w_curv * curvature), optional projection of the offset
onto the surface normal and per-point L2 clip after every step, per-sample best-result tracking and binary
search of scale_const -- restated so that it works for B > 1 and never leaves the GPU inside the loop:

  * the reference checks success with one extra victim forward PER SAMPLE and iteration (:308-330); in eval
    mode those logits are the rows of the batched forward the loss needs anyway, so ONE forward serves both;
  * best-loss / best-attack tracking and the binary search (:394-404) are torch.where updates of device
    tensors (the reference keeps Python lists and reads .item() per sample);
  * its binary-search test uses `output_label` of the LAST sample of the batch for every sample (a left-over
    loop variable, :395) -- exact for B = 1, the only batch size it was run with; here every sample uses its
    own last prediction;
  * normals come from utility.estimate_normal (k-NN select + covariance eigen-frame kernels), kappa from
    the fused kappa kernel, every adv->ori nearest-neighbour query of one iteration is one cached NN-1 sweep;
  * one whole iteration can be captured into a CUDA graph (`use_graph=True`).

`offset_proj`, `find_offset`, `lp_clip` (GeoA3_attack.py:62-101) are module functions with the reference's
signatures; `offset_proj` keeps the reference's query (it looks up the normal of the original point nearest
to the OFFSET vector itself, :68).
"""
import torch

from . import functional as F
from .knn_utils import knn_points
from .loss_utils import (_get_kappa_adv, _get_kappa_ori, chamfer_loss, curvature_loss, hausdorff_loss, norm_l2_loss,
                         pseudo_chamfer_loss)
from .utility import estimate_normal


def offset_proj(offset, ori_pc, ori_normal, project='dir'):
    """GeoA3_attack.py:62-81: project every offset onto the normal of its nearest original point (the K=1 select, then
    ONE launch for gather + normalise + project: functional.offset_proj)."""
    intra_KNN = knn_points(offset.permute(0, 2, 1), ori_pc.permute(0, 2, 1), K=1)
    return F.offset_proj(offset.contiguous(), ori_normal.contiguous(), intra_KNN.idx)


def find_offset(ori_pc, adv_pc):
    """GeoA3_attack.py:83-89: adv - (nearest original point), the K=1 select + one launch."""
    intra_KNN = knn_points(adv_pc.permute(0, 2, 1), ori_pc.permute(0, 2, 1), K=1)
    return F.find_offset(adv_pc.contiguous(), ori_pc.contiguous(), intra_KNN.idx)


def lp_clip(offset, cc_linf):
    """GeoA3_attack.py:92-101: per-point L2 clip of the offset to cc_linf (one launch, bit-identical to the torch chain)."""
    return F.lp_clip(offset.contiguous(), cc_linf)


class GeoA3Attack:
    """cfg names follow the reference's argparse options (GeoA3_attack.py:186-191)."""

    def __init__(self, model, classes, attack_method="untarget", initial_const=10., lr=0.01, optim="adam",
                 binary_max_steps=10, iter_max_steps=500, cls_loss_type="Margin", confidence=0., dis_loss_type="CD",
                 is_cd_single_side=False, dis_loss_weight=1.0, hd_loss_weight=0.1, curv_loss_weight=1.0, curv_loss_knn=16,
                 is_pro_grad=False, is_real_offset=False, cc_linf=0., is_use_lr_scheduler=False, normal_knn=3,
                 global_batch=None, use_graph=False):
        if cls_loss_type not in ("Margin", "CE", "None") or dis_loss_type not in ("CD", "L2", "None") or optim not in ("adam", "sgd"):
            raise ValueError("unsupported cls_loss_type / dis_loss_type / optim")
        if dis_loss_type == "L2" and hd_loss_weight != 0:
            raise ValueError("the reference asserts hd_loss_weight == 0 with the L2 distance loss (GeoA3_attack.py:147)")
        self.model = model.eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self.classes, self.targeted = classes, attack_method != "untarget"
        self.initial_const, self.lr, self.optim = initial_const, lr, optim
        self.binary_max_steps, self.iter_max_steps = binary_max_steps, iter_max_steps
        self.cls_loss_type, self.confidence = cls_loss_type, confidence
        self.dis_loss_type, self.is_cd_single_side = dis_loss_type, is_cd_single_side
        self.w_dis, self.w_hd, self.w_curv, self.curv_knn = dis_loss_weight, hd_loss_weight, curv_loss_weight, curv_loss_knn
        self.is_pro_grad, self.is_real_offset, self.cc_linf = is_pro_grad, is_real_offset, cc_linf
        self.use_sched, self.normal_knn = is_use_lr_scheduler, normal_knn
        self.global_batch, self.use_graph = global_batch, use_graph
        self.loop_ms = 0.0

    # GeoA3_attack.py:103-183 on [b] vectors
    def _losses(self, logits, st):
        target = st["target"]
        b = logits.shape[0]
        if self.cls_loss_type == "Margin":
            onehot = torch.zeros_like(logits).scatter_(1, target.unsqueeze(1), 1.)
            fake = (onehot * logits).sum(1)
            other = ((1. - onehot) * logits - onehot * 10000.).max(1)[0]
            cls_loss = torch.clamp((other - fake if self.targeted else fake - other) + self.confidence, min=0.)
        elif self.cls_loss_type == "CE":
            ce = torch.nn.functional.cross_entropy(logits, target, reduction="none")
            cls_loss = ce if self.targeted else -ce
        else:
            cls_loss = torch.zeros(b, device=logits.device)
        adv, ori = st["adv"], st["ori"]
        constrain = torch.zeros(b, device=logits.device)
        if self.dis_loss_type == "CD":
            constrain = constrain + self.w_dis * (pseudo_chamfer_loss(adv, ori) if self.is_cd_single_side else chamfer_loss(adv, ori))
        elif self.dis_loss_type == "L2":
            constrain = constrain + self.w_dis * norm_l2_loss(adv, ori)
        if self.w_hd != 0:
            constrain = constrain + self.w_hd * hausdorff_loss(adv, ori)
        if self.w_curv != 0:
            adv_kappa, _ = _get_kappa_adv(adv, ori, st["normal"], self.curv_knn)
            constrain = constrain + self.w_curv * curvature_loss(adv, ori, adv_kappa, st["kappa_ori"])
        return cls_loss, constrain

    def _iteration(self, st):
        ori, offset, target = st["ori"], st["offset"], st["target"]
        adv = ori + offset                                      # input_all = periodical_pc + offset (:296)
        st["adv"] = adv
        out = self.model(adv)
        logits = out[0] if isinstance(out, tuple) else out
        with torch.no_grad():                                   # :308-330, metric = constrain_loss of the PREVIOUS iteration
            pred = torch.argmax(logits, dim=1)
            ok = (pred == target) if self.targeted else (pred != target)
            metric = st["constrain"]
            better = ok & (metric < st["best_loss"])
            st["best_loss"].copy_(torch.where(better, metric, st["best_loss"]))
            st["best_attack"].copy_(torch.where(better[:, None, None], adv, st["best_attack"]))
            st["best_step"].copy_(torch.where(better, st["step"].expand_as(st["best_step"]), st["best_step"]))
            st["best_bs"].copy_(torch.where(better, st["search_step"].expand_as(st["best_bs"]), st["best_bs"]))
            it_better = ok & (metric < st["iter_best_loss"])
            st["iter_best_loss"].copy_(torch.where(it_better, metric, st["iter_best_loss"]))
            st["iter_best_score"].copy_(torch.where(it_better, pred, st["iter_best_score"]))
            st["last_ok"].copy_(ok)
        cls_loss, constrain = self._losses(logits, st)
        loss_n = cls_loss + st["scale_const"] * constrain
        denom = float(self.global_batch or loss_n.shape[0])
        loss = loss_n.sum() / denom                             # loss_n.mean() over the GLOBAL batch when sharded
        st["opt"].zero_grad(set_to_none=False)
        loss.backward()
        st["opt"].step()
        with torch.no_grad():
            st["constrain"].copy_(constrain.detach())
            st["loss_n"].copy_(loss_n.detach())
            if self.use_sched:                                  # ExponentialLR(gamma=0.9990), :292
                for gr in st["opt"].param_groups:
                    if isinstance(gr["lr"], torch.Tensor):
                        gr["lr"].mul_(0.9990)
                    else:
                        gr["lr"] = gr["lr"] * 0.9990
            if self.is_pro_grad:                                # :356-362
                if self.is_real_offset:
                    offset.copy_(find_offset(ori, ori + offset))
                offset.copy_(offset_proj(offset, ori, st["normal"]))
            if self.cc_linf != 0:                               # :364-367
                offset.copy_(lp_clip(offset, self.cc_linf))
            if self.is_pro_grad or self.cc_linf != 0:
                F.clear_cache()                                 # the offset changed behind autograd's version counter
            st["step"].add_(1)

    def attack(self, pc, label, seed=0, first_sample=0, init_offset=None):
        """pc [B,N,3], label [B] (ground truth; the target class when attack_method is targeted)
        -> (best_attack [B,3,N], success mask [B], best_loss [B], best_attack_step [B]).
        init_offset [binary_max_steps, B, 3, N] replaces the per-sample generator draws of N(0, 1e-3)
        (parity tests replay the reference's own nn.init.normal_ draws, GeoA3_attack.py:283)."""
        from .sharding import per_sample_noise
        dev = pc.device
        if dev.type != "cuda":
            raise RuntimeError("GeoA3Attack runs on CUDA only")
        B, N = pc.shape[:2]
        ori = pc.float().transpose(1, 2).contiguous().detach()
        target = label.long().to(dev).view(-1)
        with torch.no_grad():
            normal = estimate_normal(ori, k=self.normal_knn)                      # :226
            kappa_ori = _get_kappa_ori(ori, normal, self.curv_knn) if self.w_curv != 0 else None
        lower = torch.zeros(B, device=dev)
        upper = torch.full((B,), 1e10, device=dev)
        st = {
            "ori": ori, "normal": normal, "kappa_ori": kappa_ori, "target": target,
            "scale_const": torch.full((B,), float(self.initial_const), device=dev),
            "best_loss": torch.full((B,), 1e10, device=dev),
            "best_attack": torch.ones((B, 3, N), device=dev),
            "best_step": torch.full((B,), -1, dtype=torch.long, device=dev),
            "best_bs": torch.full((B,), -1, dtype=torch.long, device=dev),
            "iter_best_loss": torch.full((B,), 1e10, device=dev),
            "iter_best_score": torch.full((B,), -1, dtype=torch.long, device=dev),
            "constrain": torch.full((B,), 1e10, device=dev),
            "loss_n": torch.zeros(B, device=dev),
            "last_ok": torch.zeros(B, dtype=torch.bool, device=dev),
            "step": torch.zeros((), dtype=torch.long, device=dev),
            "search_step": torch.zeros((), dtype=torch.long, device=dev),
            "offset": torch.zeros((B, 3, N), device=dev).requires_grad_(True),
        }
        spans = []
        for search_step in range(self.binary_max_steps):
            with torch.no_grad():
                if init_offset is not None:
                    st["offset"].copy_(init_offset[search_step].to(dev))
                else:
                    st["offset"].copy_(per_sample_noise((3, N), first_sample, B, 1e-3, seed=seed + search_step, device=dev))
                st["iter_best_loss"].fill_(1e10); st["iter_best_score"].fill_(-1); st["constrain"].fill_(1e10)
                st["step"].zero_(); st["search_step"].fill_(search_step)
            st["offset"].grad = None
            lr = torch.tensor(float(self.lr), device=dev) if self.use_graph else self.lr
            if self.optim == "adam":
                st["opt"] = torch.optim.Adam([st["offset"]], lr=lr, capturable=self.use_graph, foreach=True)
            else:
                st["opt"] = torch.optim.SGD([st["offset"]], lr=float(self.lr))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if self.use_graph:
                graph = self._capture(st)
                e0.record()
                for _ in range(self.iter_max_steps):
                    graph.replay()
            else:
                e0.record()
                for _ in range(self.iter_max_steps):
                    self._iteration(st)
            e1.record()
            spans.append((e0, e1))
            with torch.no_grad():                                 # :394-404
                succ = st["last_ok"] & (st["iter_best_score"] != -1)
                sc = st["scale_const"]
                lower = torch.where(succ, torch.maximum(lower, sc), lower)
                upper = torch.where(succ, upper, torch.minimum(upper, sc))
                mid = (lower + upper) * 0.5
                sc.copy_(torch.where(upper < 1e9, mid, torch.where(succ, sc * 2, sc)))
        torch.cuda.synchronize(dev)
        self.loop_ms = sum(a.elapsed_time(b) for a, b in spans)
        return st["best_attack"], st["best_loss"] < 1e10, st["best_loss"], st["best_step"]

    def _capture(self, st, warmup=3):
        """Capture one iteration (see cw_loop.CWAttack._capture): warm-up on the real state, state restored afterwards."""
        keys = [k for k, v in st.items() if isinstance(v, torch.Tensor) and k not in ("ori", "normal", "kappa_ori", "target", "adv")]
        saved = {k: st[k].detach().clone() for k in keys}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._iteration(st)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._iteration(st)
        with torch.no_grad():
            for k, v in saved.items():
                st[k].copy_(v)
            for state in st["opt"].state.values():
                for t in state.values():
                    if isinstance(t, torch.Tensor):
                        t.zero_()
            for gr in st["opt"].param_groups:
                if isinstance(gr["lr"], torch.Tensor):
                    gr["lr"].fill_(float(self.lr))
            if st["offset"].grad is not None:
                st["offset"].grad.zero_()
        return g
