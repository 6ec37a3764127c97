"""CPU oracle for the loss-only tail of the reference's eval forward — TEST INFRASTRUCTURE ONLY.

The reference computes its training losses even in eval mode (``PtTransformer.forward``,
libs/modeling/multimodal_meta_archs.py:504-509, returns ``(results, losses)`` at :522).  This file restates
that tail literally (per-sample Python loops and all) on top of ``oracle/model_ref.py``; nothing in the
product path imports it.  SURVEY.md §8f rank 4.

Parity pin: ``tests/golden/make_golden_losses.py`` runs the REAL reference on a seeded batch with event
targets (``synth.add_event_targets``) and commits the seven values of its ``losses`` dict plus the
intermediate selections; ``tests/test_losses.py`` checks this restatement against them.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import model_ref as R

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------- elementary losses
def sigmoid_focal_loss(inputs, targets, alpha: float = 0.25, gamma: float = 2.0):
    """libs/modeling/losses.py:5-53 with reduction='sum'."""
    inputs, targets = inputs.float(), targets.float()
    p = torch.sigmoid(inputs)
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.sum()


def focal_loss_score(pred, target, alpha: float = 0.25, gamma: float = 2):
    """libs/modeling/multimodal_backbones.py:1236-1270 with reduction='sum' (log of the clamped probability)."""
    p = torch.sigmoid(pred)
    p_t = p * target + (1 - p) * (1 - target)
    a_t = alpha * target + (1 - alpha) * (1 - target)
    return (-a_t * (1 - p_t).pow(gamma) * p_t.clamp(min=1e-7).log()).sum()


def ctr_diou_loss_1d(input_offsets, target_offsets, eps: float = 1e-8):
    """libs/modeling/losses.py:56-126 with reduction='sum', class_aware=True: inputs [P, ncls, 2]; only the (point, class)
    pairs with a non-zero target offset enter (:96-99)."""
    m = torch.logical_or(target_offsets[:, :, 0] > 0, target_offsets[:, :, 1] > 0)
    io, to = input_offsets[m].float(), target_offsets[m].float()
    lp, rp, lg, rg = io[:, 0], io[:, 1], to[:, 0], to[:, 1]
    lkis, rkis = torch.min(lp, lg), torch.min(rp, rg)
    intsctk = rkis + lkis
    unionk = (lp + rp) + (lg + rg) - intsctk
    iouk = intsctk / unionk.clamp(min=eps)
    len_c = torch.max(lp, lg) + torch.max(rp, rg)
    rho = 0.5 * (rp - lp - rg + lg)
    return (1.0 - iouk + torch.square(rho / len_c.clamp(min=eps))).sum()


# ------------------------------------------------------------------------------- Alignment tail
def select_contrastive_embedding(score, embedding, mask, label, cls_prd, cls_gt, ratio: int = 8):
    """libs/modeling/multimodal_backbones.py:1080-1124.  ``mask`` arrives already shortened by one column (the caller
    passes ``mask[:, 1:]`` of the CLS-free mask, :1221), so ``length`` is the valid length minus one.  Returns the key /
    non-key embeddings and (for the tests) the selected non-key frame indices."""
    from scipy import ndimage
    B = score.shape[0]
    keys, nonkeys, picked = [], [], []
    for i in range(B):
        length = int(mask[i].long().sum())
        num = max(1.0, length / ratio)                    # true division (:1091-1092): a float count
        key_idx = label[i].bool()
        key_emb = embedding[i, key_idx]
        key_label = cls_gt[i, key_idx][0]
        grown = ndimage.binary_dilation(label[i].numpy(), iterations=4)       # +-4 frames around every key frame (:1098)
        s = F.softmax(score[i, :length], dim=-1)
        order = s.sort(descending=True)[1]
        sel: List[int] = []
        for j in order.tolist():
            if not grown[j]:
                if cls_prd[i, j] == key_label:
                    sel.append(j)
            if len(sel) >= num:
                break
        keys.append(key_emb)
        nonkeys.append(embedding[i, sel])
        picked.append(sel)
    return keys, nonkeys, picked


def alignment_tail(sd: SD, video, text, cls_video, cls_text, mask, start_end, scores_gt, m_labels, p: str = "alignment"):
    """libs/modeling/multimodal_backbones.py:1206-1233.  video / text: the Alignment outputs as [B,T,C] (before the final
    transpose); mask [B,T] bool; start_end / scores_gt [B,T]; m_labels [B,T,ncls]."""
    def score_head(x, who):       # Conv1d(C, 1, 1) over the channel axis (:1021, :1210)
        return F.conv1d(x.permute(0, 2, 1), sd[f"{p}.fc_{who}_score.weight"], sd[f"{p}.fc_{who}_score.bias"]).squeeze(1)

    def cls_head(x, who):         # Linear(C, ncls) (:1023, :1212)
        return F.linear(x, sd[f"{p}.fc_{who}_cls.weight"], sd[f"{p}.fc_{who}_cls.bias"])

    out = {}
    gt_cls = torch.argmax(m_labels, dim=2)
    for who, x, cls_tok in (("video", video, cls_video), ("text", text, cls_text)):
        sc = score_head(x, who)
        out[f"score_loss_{who}"] = focal_loss_score(sc[mask], scores_gt[mask])
        seg_cls = cls_head(x, who)
        keys, nonkeys, picked = select_contrastive_embedding(sc, x, mask[:, 1:], start_end, torch.argmax(seg_cls, dim=2), gt_cls)
        out[f"key_{who}_list"], out[f"nonkey_{who}_list"], out[f"picked_{who}"] = keys, nonkeys, picked
        out[f"cls_{who}"] = cls_tok
    return out


# ------------------------------------------------------------------------------- contrastive losses
def nce(q, k, neg, logit_scale):
    """NCE.forward (libs/modeling/multimodal_meta_archs.py:24-35); the logits are multiplied by the raw parameter
    (log(1/0.07) at init), not by its exponential (:31)."""
    q, k, neg = F.normalize(q, dim=1), F.normalize(k, dim=1), F.normalize(neg, dim=1)
    logits = torch.cat([q @ k.T, q @ neg.T], dim=1) * logit_scale
    return F.cross_entropy(logits, torch.zeros(logits.shape[0], dtype=torch.long))


def dual_contrastive_loss(sd: SD, pairs, p: str = "contrastive_losses"):
    """Dual_Contrastive_Loss.forward with reduce='sum' (libs/modeling/multimodal_meta_archs.py:48-97)."""
    cv = F.normalize(pairs["cls_video"].squeeze(1), dim=1)
    ct = F.normalize(pairs["cls_text"].squeeze(1), dim=1)
    B = cv.shape[0]
    lv = sd[p + ".logit_scale_inter"].exp() * cv @ ct.t()
    target = torch.arange(B)
    inter = (F.cross_entropy(lv, target, reduction="sum") + F.cross_entropy(lv.t(), target, reduction="sum")) / 2
    intra = 0
    for i in range(B):
        kv = pairs["key_video_list"][i].mean(0, keepdim=True)
        kt = pairs["key_text_list"][i].mean(0, keepdim=True)
        a = nce(kv, kt, pairs["nonkey_video_list"][i], sd[p + ".NCE_video.logit_scale"])
        b = nce(kt, kv, pairs["nonkey_text_list"][i], sd[p + ".NCE_text.logit_scale"])
        intra = intra + (a + b) / 2
    return inter, intra / B


# ------------------------------------------------------------------------------- PtTransformer.losses
def detection_losses(fpn_masks, logits, offsets, gt_cls_labels, gt_offsets, pairs, inter, intra, cfg: dict,
                     loss_normalizer: float):
    """PtTransformer.losses with reduce='sum' (libs/modeling/multimodal_meta_archs.py:607-686).  ``cfg``: loss_weight,
    label_smoothing, num_classes, inter_contr_weight, intra_contr_weight, score_V_weight, score_A_weight.  Returns
    (dict of the 7 losses, updated loss_normalizer).  NB the reference divides by ``B = len(fpn_masks)``, which is the number
    of pyramid LEVELS (:614), not the batch size."""
    nlev = len(fpn_masks)
    valid = torch.cat(fpn_masks, dim=1)
    gt_cls = gt_cls_labels
    pos = torch.logical_and(gt_cls.sum(-1) > 0, valid)
    pred_off = torch.cat(offsets, dim=1)[pos]
    gt_off = gt_offsets[pos]
    num_pos = int(pos.sum())
    loss_normalizer = 0.9 * loss_normalizer + 0.1 * max(num_pos, 1)                      # (:637-640)
    tgt = gt_cls[valid]
    tgt = tgt * (1 - cfg["label_smoothing"]) + cfg["label_smoothing"] / (cfg["num_classes"] + 1)
    cls_loss = sigmoid_focal_loss(torch.cat(logits, dim=1)[valid], tgt) / loss_normalizer
    if num_pos == 0:
        reg_loss = 0 * pred_off.sum()
    else:
        reg_loss = ctr_diou_loss_1d(pred_off, gt_off) / loss_normalizer
    lw = cfg["loss_weight"] if cfg["loss_weight"] > 0 else cls_loss / max(float(reg_loss), 0.01)
    sv, st = pairs["score_loss_video"], pairs["score_loss_text"]
    final = (cls_loss + reg_loss * lw + inter * cfg["inter_contr_weight"] + intra * cfg["intra_contr_weight"]
             + sv * cfg["score_V_weight"] + st * cfg["score_A_weight"])
    out = {"cls_loss": cls_loss / nlev, "reg_loss": reg_loss * lw / nlev, "final_loss": final / nlev,
           "inter_contr_loss": inter * cfg["inter_contr_weight"] / nlev,
           "intra_contr_loss": intra * cfg["intra_contr_weight"] / nlev,
           "score_loss_video": sv * cfg["score_V_weight"] / nlev, "score_loss_audio": st * cfg["score_A_weight"] / nlev}
    return out, loss_normalizer


def forward_losses(sd: SD, batch: dict, cfg: dict, loss_normalizer: float, pos_embd=None):
    """The whole eval forward of the reference up to its ``losses`` dict (multimodal_meta_archs.py:426-509) on the oracle.
    Returns (losses, loss_normalizer, intermediates)."""
    sd = dict(sd)
    visual, audio, mask = batch["visual"], batch["audio"], batch["mask"]
    if "backbone.pos_embd" not in sd:
        C = sd["backbone.embd_V.0.conv.weight"].shape[0]
        sd["backbone.pos_embd"] = pos_embd if pos_embd is not None else R.sinusoid_pos_embd(visual.shape[-1], C)
    v, a, cls_v, cls_t = R.alignment(sd, "alignment", visual, audio, mask, return_cls=True)
    pairs = alignment_tail(sd, v.transpose(1, 2), a.transpose(1, 2), cls_v, cls_t, mask[:, 0], batch["start_end"],
                           batch["scores"], batch["m_labels"])
    fv, fa, ms = R.backbone(sd, "backbone", v, a, mask)
    feats = [torch.cat((x, y), 1) for x, y in zip(fv, fa)]
    logits, offsets = R.heads(sd, feats, ms)
    inter, intra = dual_contrastive_loss(sd, pairs)
    losses, ln = detection_losses([m.squeeze(1) for m in ms], logits, offsets, batch["gt_cls_labels"], batch["gt_offsets"],
                                  pairs, inter, intra, cfg, loss_normalizer)
    inter_d = {"video": v.transpose(1, 2), "text": a.transpose(1, 2), "cls_video": cls_v, "cls_text": cls_t,
               "logits": logits, "offsets": offsets, "masks": [m.squeeze(1) for m in ms], "pairs": pairs}
    return losses, ln, inter_d
