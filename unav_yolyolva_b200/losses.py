"""The loss-only tail of the reference's eval forward, device resident and free of host synchronisation
(SURVEY.md §8f rank 4).

The reference computes its seven training losses even in eval mode (``PtTransformer.forward`` calls ``self.losses`` at
/root/reference/libs/modeling/multimodal_meta_archs.py:504-509 and returns ``(results, losses)`` at :522); ``valid_one_epoch``
only averages them (libs/utils/train_utils.py:409-415).  None of it influences the detections, so it is kept off the fused
engine: after the engine's forward of a batch, ``eval_losses`` reads what the engine left in the plan's buffers —

* ``P["Y"]``        Alignment's ``Linear -> ReLU`` output before its last LayerNorm (multimodal_backbones.py:1197-1198),
* ``P["F"]``        the token matrix after the multiway layers, whose first row per item is the [CLSV] / [CLST] token (:1189-1190),
* ``P["logits"]`` / ``P["offsets"]`` / ``P["m_heads"]``   the head outputs, 441 rows per video,

and evaluates the losses with batched torch ops on the device (SURVEY.md §8b allows the loss code to stay torch).  What the
reference does with per-sample Python loops, ``.item()`` calls, boolean indexing and a scipy ``binary_dilation`` on the
host (multimodal_backbones.py:1080-1124, multimodal_meta_archs.py:75-90, :617-640) is restated here with masks:
no host round trip, no data-dependent shapes, so it can follow the CUDA graph on the same stream.

Differences, all outside the returned numbers: ``loss_normalizer`` is kept as a 0-d FP64 device tensor (the reference keeps
a Python float and synchronises on ``num_pos.item()``); a video without any ``start_end`` frame raises an ``IndexError`` in
the reference (:1096) and contributes an all-zero key embedding here.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

LOSS_KEYS = ("cls_loss", "reg_loss", "final_loss", "inter_contr_loss", "intra_contr_loss", "score_loss_video",
             "score_loss_audio")
GT_KEYS = ("scores", "start_end", "m_labels", "gt_offsets", "gt_cls_labels")


def _focal_sigmoid_sum(x, t, w, alpha: float = 0.25, gamma: float = 2.0):
    """sigmoid_focal_loss(reduction='sum') (libs/modeling/losses.py:36-51) over the elements where ``w`` is set."""
    p = torch.sigmoid(x)
    ce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    p_t = p * t + (1 - p) * (1 - t)
    loss = (alpha * t + (1 - alpha) * (1 - t)) * ce * (1 - p_t) ** gamma
    return torch.where(w, loss, torch.zeros_like(loss)).sum()


def _focal_score_sum(x, t, w, alpha: float = 0.25, gamma: float = 2.0):
    """focal_loss_score(reduction='sum') (multimodal_backbones.py:1254-1259) over the elements where ``w`` is set."""
    p = torch.sigmoid(x)
    p_t = p * t + (1 - p) * (1 - t)
    a_t = alpha * t + (1 - alpha) * (1 - t)
    fl = -a_t * (1 - p_t).pow(gamma) * p_t.clamp(min=1e-7).log()
    return torch.where(w, fl, torch.zeros_like(fl)).sum()


def select_nonkey(score, mask_short, start_end, cls_prd, cls_gt, ratio: int = 8):
    """Mask form of ``Alignment.select_contrastive_embedding`` (multimodal_backbones.py:1080-1124).

    score [B,T] frame scores, mask_short [B,T-1] (the caller's ``mask[:, 1:]``, :1221), start_end [B,T], cls_prd [B,T]
    predicted class per frame, cls_gt [B,T] class of ``m_labels`` per frame.  Returns (key [B,T] bool, nonkey [B,T] bool):
    the reference walks the frames of ``score[:length]`` by descending score and keeps those outside the event dilated by 4
    frames whose predicted class equals the event's class, until ``len >= max(1, length / 8)`` (a float: ceil)."""
    B, T = score.shape
    pos = torch.arange(T, device=score.device)[None]
    length = mask_short.sum(1)
    need = torch.ceil(torch.clamp(length.to(torch.float32) / ratio, min=1.0))
    key = start_end != 0
    first = torch.where(key, pos, torch.full_like(pos, T - 1)).amin(1, keepdim=True)
    key_label = cls_gt.gather(1, first)
    grown = F.max_pool1d(key.to(torch.float32)[:, None], 9, 1, 4)[:, 0] > 0                    # binary_dilation x4 (:1098)
    in_len = pos < length[:, None]
    order = torch.where(in_len, score, torch.full_like(score, float("-inf"))).argsort(dim=1, descending=True, stable=True)
    elig = (~grown & (cls_prd == key_label) & in_len).gather(1, order)
    take = elig & (elig.cumsum(1) <= need[:, None])
    nonkey = torch.zeros_like(take).scatter(1, order, take)
    return key, nonkey


def _nce(q, k, emb_n, nonkey, logit_scale):
    """NCE.forward for every sample at once (multimodal_meta_archs.py:24-35): q, k [B,C] key means, emb_n [B,T,C] normalised
    frame embeddings, nonkey [B,T] the negatives of each sample.  Cross entropy against index 0 of [l_pos | l_neg]."""
    q, k = F.normalize(q, dim=1), F.normalize(k, dim=1)
    l_pos = (q * k).sum(1, keepdim=True) * logit_scale
    l_neg = torch.einsum("bc,btc->bt", q, emb_n) * logit_scale
    l_neg = torch.where(nonkey, l_neg, torch.full_like(l_neg, float("-inf")))
    return torch.logsumexp(torch.cat((l_pos, l_neg), 1), dim=1) - l_pos[:, 0]


@torch.no_grad()
def losses_from_activations(model, video, text, cls_video, cls_text, logits, offsets, valid, video_list) -> Dict[str, torch.Tensor]:
    """video / text [B,T,C] Alignment outputs, cls_* [B,C], logits [B,Ttot,ncls], offsets [B,Ttot,ncls,2], valid [B,Ttot] bool,
    ``video_list`` the collate dict with the GT tensors (any device).  Updates ``model.loss_normalizer``."""
    dev = video.device
    al, cl = model.alignment, model.contrastive_losses
    mask = video_list["mask"].to(dev).reshape(video.shape[0], -1).bool()
    scores_gt = video_list["scores"].to(dev, torch.float32)
    start_end = video_list["start_end"].to(dev, torch.float32)
    m_labels = video_list["m_labels"].to(dev, torch.float32)
    gt_cls = video_list["gt_cls_labels"].to(dev, torch.float32)
    gt_off = video_list["gt_offsets"].to(dev, torch.float32)
    B = video.shape[0]

    # ---- Alignment tail (multimodal_backbones.py:1209-1233)
    cls_gt = m_labels.argmax(2)
    sel, score_loss = {}, {}
    for who, x in (("video", video), ("text", text)):
        sc_mod, cls_mod = getattr(al, f"fc_{who}_score"), getattr(al, f"fc_{who}_cls")
        sc = F.linear(x, sc_mod.weight.reshape(1, -1).float(), sc_mod.bias.float())[..., 0]        # Conv1d(C,1,1) (:1210)
        score_loss[who] = _focal_score_sum(sc, scores_gt, mask)
        prd = F.linear(x, cls_mod.weight.float(), cls_mod.bias.float()).argmax(2)
        sel[who] = select_nonkey(sc, mask[:, 1:], start_end, prd, cls_gt)

    # ---- Dual_Contrastive_Loss, reduce='sum' (multimodal_meta_archs.py:48-97)
    cv, ct = F.normalize(cls_video, dim=1), F.normalize(cls_text, dim=1)
    lv = cl.logit_scale_inter.float().exp() * cv @ ct.t()
    tgt = torch.arange(B, device=dev)
    inter = (F.cross_entropy(lv, tgt, reduction="sum") + F.cross_entropy(lv.t(), tgt, reduction="sum")) / 2
    key = sel["video"][0]
    kcnt = key.sum(1, keepdim=True).clamp(min=1).to(torch.float32)
    kv = (video * key[..., None]).sum(1) / kcnt
    kt = (text * key[..., None]).sum(1) / kcnt
    a = _nce(kv, kt, F.normalize(video, dim=2), sel["video"][1], cl.NCE_video.logit_scale.float())
    b = _nce(kt, kv, F.normalize(text, dim=2), sel["text"][1], cl.NCE_text.logit_scale.float())
    intra = ((a + b) / 2).sum() / B

    # ---- PtTransformer.losses, reduce='sum' (multimodal_meta_archs.py:607-686)
    nlev = len(model.fpn_strides)                    # the reference's `B = len(fpn_masks)`: pyramid levels (:614)
    pos = (gt_cls.sum(-1) > 0) & valid
    num_pos = pos.sum().to(torch.float64)
    ln = model.loss_normalizer
    if not torch.is_tensor(ln) or ln.device != dev:
        ln = torch.as_tensor(float(ln), dtype=torch.float64, device=dev)
    mom = model.loss_normalizer_momentum
    ln = mom * ln + (1 - mom) * num_pos.clamp(min=1)
    model.loss_normalizer = ln
    ls = model.train_label_smoothing
    tgt_cls = gt_cls * (1 - ls) + ls / (model.num_classes + 1)
    cls_loss = (_focal_sigmoid_sum(logits, tgt_cls, valid[..., None].expand_as(logits)) / ln).to(torch.float32)
    # ctr_diou_loss_1d, class aware: (point, class) pairs of positive points with a non-zero target (losses.py:96-124)
    m = pos[..., None] & ((gt_off[..., 0] > 0) | (gt_off[..., 1] > 0))
    lp, rp, lg, rg = offsets[..., 0], offsets[..., 1], gt_off[..., 0], gt_off[..., 1]
    inter_k = torch.min(rp, rg) + torch.min(lp, lg)
    union_k = (lp + rp) + (lg + rg) - inter_k
    len_c = torch.max(lp, lg) + torch.max(rp, rg)
    rho = 0.5 * (rp - lp - rg + lg)
    diou = 1.0 - inter_k / union_k.clamp(min=1e-8) + torch.square(rho / len_c.clamp(min=1e-8))
    reg_loss = (torch.where(m, diou, torch.zeros_like(diou)).sum() / ln).to(torch.float32)
    if model.train_loss_weight > 0:
        lw = model.train_loss_weight
    else:
        lw = cls_loss / reg_loss.clamp(min=0.01)
    sv, st = score_loss["video"], score_loss["text"]
    final = (cls_loss + reg_loss * lw + inter * model.inter_contr_weight + intra * model.intra_contr_weight
             + sv * model.score_V_weight + st * model.score_T_weight)
    return {"cls_loss": cls_loss / nlev, "reg_loss": reg_loss * lw / nlev, "final_loss": final / nlev,
            "inter_contr_loss": inter * model.inter_contr_weight / nlev,
            "intra_contr_loss": intra * model.intra_contr_weight / nlev,
            "score_loss_video": sv * model.score_V_weight / nlev, "score_loss_audio": st * model.score_T_weight / nlev}


@torch.no_grad()
def eval_losses(model, P: dict, video_list) -> Dict[str, torch.Tensor]:
    """The reference's ``losses`` dict for the batch whose forward the engine has just enqueued into plan ``P`` (same stream)."""
    eng = model.engine
    B, T, C = P["B"], eng.T, eng.C
    half = B * T
    al = model.alignment
    outs = []
    for g, who in enumerate(("video", "text")):          # the last LayerNorm of fc_video / fc_text (:1197-1198)
        ln = getattr(al, f"fc_{who}")[3]
        outs.append(F.layer_norm(P["Y"][g * half:(g + 1) * half].view(B, T, C), (C,), ln.weight.float(), ln.bias.float(), ln.eps))
    tok = P["F"].view(2 * B, T + 1, C)
    logits = P["logits"].view(B, eng.Ttot, eng.ncls)
    offsets = P["offsets"].view(B, eng.Ttot, eng.ncls, 2)
    valid = P["m_heads"].view(B, eng.Ttot).bool()
    return losses_from_activations(model, outs[0], outs[1], tok[:B, 0], tok[B:, 0], logits, offsets, valid, video_list)
