"""``LocPointTransformer`` meta-architecture and its heads, with the reference's constructor, parameter
names and eval-time calling convention (/root/reference/libs/modeling/multimodal_meta_archs.py:101-875):

    results, losses = model(video_list)          # video_list = collate dict (libs/datasets/data_utils.py:214-229)
    results = {'segments': [B,K,2] f32 seconds, 'scores': [B,K] f32, 'labels': [B,K] i64}

``forward`` runs the fused sm_100a engine (``..engine.HotPathEngine``): Alignment -> backbone -> heads ->
decode -> per-class soft-NMS -> seconds, all device resident; nothing is computed by PyTorch eager ops and
there is no CPU fallback.  Deviations from the reference, all outside the detection results:

* ``losses``: the reference computes the training losses even in eval (:504-509).  They are loss-only work
  (SURVEY.md §8f rank 4): when the collate dict carries the GT tensors and ``model.eval_losses`` is True (default) they are
  evaluated on the device behind the engine's forward (``..losses.eval_losses``: batched torch ops, no host
  synchronisation) and returned under the reference's 7 keys; otherwise (``eval_losses = False``, a dict without GT, the
  asynchronous ``submit`` path) the 7 keys hold zero scalars.
* a video with fewer than ``max_seg_num`` detections makes the reference's ``torch.cat`` raise (:869-873); the
  same condition raises a ``RuntimeError`` here (the padded device outputs stay available through
  ``model.engine``).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
from torch import nn

from .blocks import LayerNorm, MaskedConv1D, Scale
from .models import make_dependency_block, make_multimodal_backbone, register_multimodal_meta_arch
from .multimodal_backbones import Alignment
from .. import losses as _losses
from ..losses import GT_KEYS, LOSS_KEYS


class NCE(nn.Module):
    """Loss-only parameter holder (:19-35); evaluated by ``..losses._nce``."""

    def __init__(self):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))


class Dual_Contrastive_Loss(nn.Module):
    """Loss-only parameter holder (:37-98); evaluated by ``..losses.losses_from_activations``."""

    def __init__(self, args=None):
        super().__init__()
        self.logit_scale_inter = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.NCE_video = NCE()
        self.NCE_text = NCE()


def _head_trunk(owner, input_dim, feat_dim, num_layers, kernel_size, with_ln):
    owner.head, owner.norm = nn.ModuleList(), nn.ModuleList()
    for idx in range(num_layers - 1):
        owner.head.append(MaskedConv1D(input_dim if idx == 0 else feat_dim, feat_dim, kernel_size, stride=1,
                                       padding=kernel_size // 2, bias=not with_ln))
        owner.norm.append(LayerNorm(feat_dim) if with_ln else nn.Identity())


class PtTransformerClsHead(nn.Module):
    """Shared 1-D conv classification head (:101-178)."""

    def __init__(self, input_dim, feat_dim, num_classes, prior_prob=0.01, num_layers=3, kernel_size=3,
                 act_layer=nn.ReLU, with_ln=False, empty_cls=()):
        super().__init__()
        self.act = act_layer()
        _head_trunk(self, input_dim, feat_dim, num_layers, kernel_size, with_ln)
        self.cls_head = MaskedConv1D(feat_dim, num_classes, kernel_size, stride=1, padding=kernel_size // 2)
        nn.init.constant_(self.cls_head.conv.bias, -math.log((1 - prior_prob) / prior_prob))
        for idx in empty_cls:
            nn.init.constant_(self.cls_head.conv.bias[idx], -math.log((1 - 1e-6) / 1e-6))

    def forward(self, fpn_feats, fpn_masks):
        from .. import _fwd
        return _fwd.cls_head_forward(self, fpn_feats, fpn_masks)


class PtTransformerRegHead(nn.Module):
    """Shared 1-D conv regression head with a per-level ``Scale`` (:181-259)."""

    def __init__(self, input_dim, feat_dim, num_classes, fpn_levels, num_layers=3, kernel_size=3, act_layer=nn.ReLU,
                 with_ln=False, class_aware=False):
        super().__init__()
        self.fpn_levels, self.num_classes = fpn_levels, num_classes
        self.act = act_layer()
        _head_trunk(self, input_dim, feat_dim, num_layers, kernel_size, with_ln)
        self.scale = nn.ModuleList(Scale() for _ in range(fpn_levels))
        self.offset_head = MaskedConv1D(feat_dim, 2 * num_classes if class_aware else 2, kernel_size, stride=1,
                                        padding=kernel_size // 2)

    def forward(self, fpn_feats, fpn_masks):
        from .. import _fwd
        return _fwd.reg_head_forward(self, fpn_feats, fpn_masks)


class PendingDetections:
    """Handle returned by ``PtTransformer.submit``: detections of one step, landing in pinned host memory."""

    def __init__(self, slot, event, losses=None, plan=None):
        self._slot, self._event = slot, event
        # the engine plan of this step: device consumers may chain work on ``plan["nms_stream"]`` and read ``plan["out_*"]``
        # there (valid until the plan's next step, which is enqueued behind it on the same stream)
        self.plan = plan
        self.losses = losses        # the reference's `losses` dict (0-d device tensors) if requested at submit(), else None

    def done(self) -> bool:
        return self._event.query()

    def result(self):
        """Host tensors {segments [B,K,2], scores [B,K], labels [B,K]} (views of the staging slot: consume or clone
        them before submitting two more steps).  Same ragged-count error as ``forward``."""
        self._event.synchronize()
        counts = self._slot["counts"]
        K_ = int(counts.max()) if counts.numel() else 0
        if int(counts.min()) != K_:
            raise RuntimeError("videos in the batch produced different numbers of detections "
                               f"({counts.tolist()}); the reference's torch.cat fails the same way "
                               "(multimodal_meta_archs.py:869-873)")
        return {"segments": self._slot["segments"][:, :K_], "scores": self._slot["scores"][:, :K_],
                "labels": self._slot["labels"][:, :K_]}


@register_multimodal_meta_arch("LocPointTransformer")
class PtTransformer(nn.Module):
    """Single-stage audio-visual event localiser; constructor identical to the reference (:267-295)."""

    # precision mode of the fused engine: "bf16x3" (tcgen05, FP32-accurate split), "bf16" (tcgen05), "fp32" (FFMA)
    precision = "bf16x3"
    # engine plans (sets of buffers + streams) that submit() cycles through: 3 batches in flight (measured on B200, batch
    # 16: 4.9 ms/step with one, 3.87 with two, 3.67 with three, 3.66 with four)
    streams = 3
    use_cuda_graph = True
    # evaluate the reference's (loss-only) `losses` dict in forward() when the batch carries GT tensors
    eval_losses = True

    def __init__(self, backbone_type, dependency_type, backbone_arch, scale_factor, input_dim_V, input_dim_A,
                 max_seq_len, n_head, embd_kernel_size, embd_dim, embd_with_ln, head_dim, regression_range,
                 head_num_layers, head_kernel_size, head_with_ln, use_abs_pe, num_classes, train_cfg, test_cfg,
                 class_aware, use_dependency, intra_contr_weight, inter_contr_weight, score_V_weight,
                 score_A_weight):
        super().__init__()
        self.fpn_strides = [scale_factor ** i for i in range(backbone_arch[-1] + 1)]
        self.reg_range = regression_range
        assert len(self.fpn_strides) == len(self.reg_range)
        self.scale_factor, self.num_classes = scale_factor, num_classes
        self.class_aware, self.use_dependency = class_aware, use_dependency
        self.max_seq_len = max_seq_len
        for stride in self.fpn_strides:
            assert max_seq_len % stride == 0, "max_seq_len must be divisible by fpn stride"
        self.max_div_factor = max(self.fpn_strides)
        self.train_loss_weight = train_cfg["loss_weight"]
        self.inter_contr_weight, self.intra_contr_weight = inter_contr_weight, intra_contr_weight
        self.score_V_weight, self.score_T_weight = score_V_weight, score_A_weight
        self.train_cls_prior_prob = train_cfg["cls_prior_prob"]
        self.train_dropout, self.train_droppath = train_cfg["dropout"], train_cfg["droppath"]
        self.train_label_smoothing = train_cfg["label_smoothing"]
        self.test_pre_nms_thresh = test_cfg["pre_nms_thresh"]
        self.test_pre_nms_topk = test_cfg["pre_nms_topk"]
        self.test_iou_threshold = test_cfg["iou_threshold"]
        self.test_min_score = test_cfg["min_score"]
        self.test_max_seg_num = test_cfg["max_seg_num"]
        self.test_nms_method = test_cfg["nms_method"]
        assert self.test_nms_method in ["soft", "hard", "none"]
        self.test_duration_thresh = test_cfg["duration_thresh"]
        self.test_multiclass_nms = test_cfg["multiclass_nms"]
        self.test_nms_sigma = test_cfg["nms_sigma"]
        self.test_voting_thresh = test_cfg["voting_thresh"]

        assert backbone_type in ["convTransformer"]
        self.backbone = make_multimodal_backbone(
            "convTransformer", n_in_V=input_dim_V, n_in_A=input_dim_A, n_embd=embd_dim, n_head=n_head,
            n_embd_ks=embd_kernel_size, max_len=max_seq_len, arch=backbone_arch, scale_factor=scale_factor,
            with_ln=embd_with_ln, attn_pdrop=0.0, proj_pdrop=self.train_dropout, path_pdrop=self.train_droppath,
            use_abs_pe=use_abs_pe)
        assert dependency_type in ["DependencyBlock"]
        if self.use_dependency:
            self.dependency_block = make_dependency_block(
                "DependencyBlock", in_channel=embd_dim * 2, n_embd=128, n_embd_ks=embd_kernel_size,
                num_classes=num_classes, path_pdrop=self.train_droppath)
        self.cls_head = PtTransformerClsHead(embd_dim * 2, head_dim, num_classes, kernel_size=head_kernel_size,
                                             prior_prob=self.train_cls_prior_prob, with_ln=head_with_ln,
                                             num_layers=head_num_layers, empty_cls=train_cfg["head_empty_cls"])
        self.reg_head = PtTransformerRegHead(embd_dim * 2, head_dim, num_classes, len(self.fpn_strides),
                                             kernel_size=head_kernel_size, num_layers=head_num_layers,
                                             with_ln=head_with_ln, class_aware=class_aware)
        self.loss_normalizer = train_cfg["init_loss_norm"]
        self.loss_normalizer_momentum = 0.9
        # raw feature dims are hard-coded by the reference (:406-409)
        self.alignment = Alignment(video_dim=2048, audio_dim=128)
        self.contrastive_losses = Dual_Contrastive_Loss()
        self._engine = None
        self._engine_key = None
        self._host_ring = {}
        self._submit_n = 0
        self._profile = None                    # dict: run_shard(stats=...) collects host-side waits here

    @property
    def device(self):
        return next(self.parameters()).device

    # ---- engine management -------------------------------------------------------------------
    def invalidate_engine(self):
        """Drop the packed weights / launch plans (call after modifying parameters in place)."""
        self._engine = None

    def load_state_dict(self, *args, **kwargs):
        self._engine = None
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):      # .to() / .cuda() / .float(): parameters move, repack lazily
        self._engine = None
        return super()._apply(fn, *args, **kwargs)

    @property
    def engine(self):
        """The fused engine, (re)built lazily: weights are packed on first use after ``load_state_dict`` /
        ``.to(device)`` (DataParallel moves the module after construction, eval.py:61)."""
        from ..engine import HotPathEngine
        key = (str(self.device), self.precision, self.use_cuda_graph)
        if self._engine is None or self._engine_key != key:
            self._engine = HotPathEngine(self, mode=self.precision, use_graph=self.use_cuda_graph)
            self._engine_key = key
        return self._engine

    # ---- forward -----------------------------------------------------------------------------
    def forward(self, video_list):
        if self.training:
            raise NotImplementedError("training is outside the inference hot path (SURVEY.md §2 C16)")
        plan = self.run_hot_path(video_list)
        if self.eval_losses and all(k in video_list for k in GT_KEYS):
            if not self.class_aware:
                raise NotImplementedError("eval losses are implemented for class_aware=True (both reference configs)")
            losses = _losses.eval_losses(self, plan, video_list)          # same stream, behind the forward; before the host sync below
        else:
            dev = self.device
            losses = {k: torch.zeros((), device=dev) for k in LOSS_KEYS}
        results = self.collect_results(plan)
        return results, losses

    def _host_slot(self, B):
        """Pinned host staging (video metadata in, detections out): a ring of ``streams + 2`` slots per batch size, so that
        ``streams + 1`` handles can be outstanding."""
        ring = self._host_ring.get(B)
        nslots = max(1, self.streams) + 2
        if ring is not None and len(ring["slots"]) != nslots:
            ring = None
        if ring is None:
            K_ = self.test_max_seg_num
            ring = {"next": 0, "slots": [
                {"meta": torch.empty(B, 4, dtype=torch.float32).pin_memory(),
                 "segments": torch.empty(B, K_, 2, dtype=torch.float32).pin_memory(),
                 "scores": torch.empty(B, K_, dtype=torch.float32).pin_memory(),
                 "labels": torch.empty(B, K_, dtype=torch.int64).pin_memory(),
                 "counts": torch.empty(B, dtype=torch.int32).pin_memory(), "event": None} for _ in range(nslots)]}
            self._host_ring[B] = ring
        slot = ring["slots"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % nslots
        if slot["event"] is not None:
            if self._profile is not None:
                import time
                t0 = time.perf_counter()
                slot["event"].synchronize()
                self._profile["host_slot_wait_s"] = self._profile.get("host_slot_wait_s", 0.0) + time.perf_counter() - t0
            else:
                slot["event"].synchronize()     # its previous user (a full ring ago) must have retired
        return slot

    @torch.no_grad()
    def run_hot_path(self, video_list, _slot=None, _overlap_nms=False, _plan_slot=0):
        """Launch the whole device-resident path for one collate dict; returns the engine plan (outputs stay
        on the device, nothing is synchronised)."""
        vis, aud, mask = video_list["visual"], video_list["audio"], video_list["mask"]
        B = vis.shape[0]
        slot = _slot if _slot is not None else self._host_slot(B)
        m = slot["meta"]                        # one host-side fill (64 scalar tensor writes cost ~0.1 ms per step)
        m.copy_(torch.tensor([[float(x) for x in video_list[k]] for k in ("feat_stride", "feat_num_frames", "fps", "duration")],
                             dtype=torch.float32).t())
        plan = self.engine.run(vis, aud, mask, m, overlap_nms=_overlap_nms, slot=_plan_slot)
        if _slot is None:                       # the slot's meta is in flight until this point of the stream
            ev = torch.cuda.Event(); ev.record(); slot["event"] = ev
        return plan

    @torch.no_grad()
    def submit(self, video_list, with_losses: bool = False):
        """Asynchronous form of ``forward`` for evaluation loops: enqueue the step and the device->host copy of its
        detections into pinned memory, return immediately.  ``handle.result()`` waits for that step only, so the host
        work of step j+1 (collate, upload, launch) overlaps the device work of step j:

            prev = None
            for batch in CudaPrefetcher(loader, device):
                cur = model.submit(batch)
                if prev is not None: consume(prev.result())
                prev = cur

        Up to ``streams + 1`` handles may be outstanding (the staging buffers are a ring of ``streams + 2``); keeping
        ``streams`` steps in flight before consuming the oldest result gives the best throughput.
        ``with_losses=True`` also evaluates the reference's loss dict for this batch (``..losses.eval_losses``) on the plan's
        forward stream, behind the decode; ``handle.losses`` then holds 0-d device tensors, ordered after ``handle.result()``
        only through a device synchronisation or ``.item()`` on them."""
        if self.training:
            raise NotImplementedError("training is outside the inference hot path (SURVEY.md §2 C16)")
        B = video_list["visual"].shape[0]
        slot = self._host_slot(B)
        # two engine plans are used alternately, each with its own streams: two steps are in flight, and the soft-NMS of a
        # step runs on its plan's side stream; the copies of its outputs to the host follow it there
        self._submit_n += 1
        plan = self.run_hot_path(video_list, _slot=slot, _overlap_nms=True, _plan_slot=self._submit_n % max(1, self.streams))
        losses = None
        if with_losses:
            if not self.class_aware:
                raise NotImplementedError("eval losses are implemented for class_aware=True (both reference configs)")
            with torch.cuda.stream(plan["fwd_stream"]):
                losses = _losses.eval_losses(self, plan, video_list)
        ns = plan["nms_stream"]
        with torch.cuda.stream(ns):
            slot["segments"].copy_(plan["out_segs"], non_blocking=True)
            slot["scores"].copy_(plan["out_scores"], non_blocking=True)
            slot["labels"].copy_(plan["out_labels"], non_blocking=True)
            slot["counts"].copy_(plan["out_counts"], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(ns); slot["event"] = ev
        return PendingDetections(slot, ev, losses, plan)

    def collect_results(self, plan):
        counts = plan["out_counts"].cpu()
        K_ = int(counts.max().item()) if counts.numel() else 0
        if int(counts.min().item()) != K_:
            raise RuntimeError("videos in the batch produced different numbers of detections "
                               f"({counts.tolist()}); the reference's torch.cat fails the same way "
                               "(multimodal_meta_archs.py:869-873)")
        return {"segments": plan["out_segs"][:, :K_].clone(), "scores": plan["out_scores"][:, :K_].clone(),
                "labels": plan["out_labels"][:, :K_].clone()}

    @torch.no_grad()
    def inference(self, video_list, fpn_masks, out_cls_logits, out_offsets):
        """Reference entry (:689-742): decode + NMS from per-level head outputs (lists of [B,T_l,C] tensors)."""
        from .. import _fwd
        return _fwd.inference_from_heads(self, video_list, fpn_masks, out_cls_logits, out_offsets)
