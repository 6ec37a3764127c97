"""Seeded synthetic weights and batches for parity tests and ``bench.py``.

There is no network for datasets or checkpoints, so both sides of every comparison use:

* ``trained_like_state_dict``: a deterministic, *name-keyed* weight generator.  Each entry of the
  state_dict manifest (``tests/golden/state_dict_manifest.json`` — the 1235 names/shapes of the
  reference's ``PtTransformer.state_dict()``) is drawn from a ``torch.Generator`` seeded with
  crc32(canonical name), so the reference model (golden generation), the oracle and the CUDA
  engine all see bit-identical weights without a 390 MB checkpoint.  Values are "trained-like":
  the reference initialises ``AffineDropPath.scale`` to 1e-4 and all biases to 0
  (/root/reference/libs/modeling/blocks.py:381-386, multimodal_backbones.py:765-769), which would
  hide whole branches from a parity test, so scales / LayerNorm affine / biases are O(1) here.
* ``make_batch``: the collate dict of /root/reference/libs/datasets/data_utils.py:214-229 with
  I3D-like (post-ReLU) visual and VGGish-like audio features, per-video seed 1234 + video index.
"""
from __future__ import annotations

import json
import math
import os
import re
import zlib
from typing import Dict, List, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
MANIFEST = os.path.join(os.path.dirname(_HERE), "tests", "golden", "state_dict_manifest.json")

# aliased modules in the reference share one tensor under several names
_ALIASES = [
    (re.compile(r"alignment\.multiway_list\.\d+\."), "alignment.multiway_list.0."),
    (re.compile(r"fusion_module\.downsample_layers\.\d+\."), "fusion_module.downsample_layers.0."),
]


def canonical_name(name: str) -> str:
    for pat, rep in _ALIASES:
        name = pat.sub(rep, name)
    return name


def load_manifest(path: str = MANIFEST) -> Dict[str, List[int]]:
    with open(path) as f:
        return json.load(f)


def _gen(name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(zlib.crc32(canonical_name(name).encode()) & 0x7FFFFFFF)
    return g


def trained_like_tensor(name: str, shape: List[int]) -> torch.Tensor:
    g = _gen(name)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]

    def normal(std, mean=0.0):
        return torch.randn(shape, generator=g) * std + mean

    def uniform(lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    if "logit_scale" in name:
        return torch.full(shape, math.log(1 / 0.07))
    if leaf == "scale":                                   # AffineDropPath.scale / Scale.scale
        return uniform(0.5, 1.5)
    if name.endswith("attn_block.bias"):
        return normal(0.5)
    if "pos_embed_" in name or "type_" in name or "cls_token_" in name:
        return normal(0.1)
    is_channel_ln = len(shape) == 3 and shape[0] == 1 and shape[2] == 1
    is_nn_ln = len(shape) == 1 and (
        re.search(r"\.norm\w*\.(weight|bias)$", name) is not None
        or re.search(r"alignment\.fc_(video|text)\.3\.", name) is not None)
    if is_channel_ln or is_nn_ln:
        return uniform(0.5, 1.5) if leaf == "weight" else normal(0.1)
    if leaf == "bias":
        if name == "cls_head.cls_head.conv.bias":
            return normal(1.0, mean=-4.6)                # prior-prob bias -log(99) +- 1
        return normal(0.1)
    if leaf == "weight" and len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return normal(1.0 / math.sqrt(fan_in))
    raise ValueError(f"no generation rule for {name} {shape}")


def trained_like_state_dict(manifest: Optional[Dict[str, List[int]]] = None,
                            prefix: str = "") -> Dict[str, torch.Tensor]:
    manifest = manifest or load_manifest()
    cache: Dict[str, torch.Tensor] = {}
    out = {}
    for name, shape in manifest.items():
        cn = canonical_name(name)
        if cn not in cache:
            cache[cn] = trained_like_tensor(cn, shape)
        out[prefix + name] = cache[cn]
    return out


# ------------------------------------------------------------------------------- batches
def video_length(index: int, lo: int = 60, hi: int = 187) -> int:
    g = torch.Generator().manual_seed(1234 + index)
    return int(torch.randint(lo, hi + 1, (1,), generator=g).item())


def make_points(T: int, n_levels: int = 6, scale_factor: int = 2,
                regression_range=((0, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 10000))):
    """Per-level [T_l, 4] = (t, reg_lo, reg_hi, stride): /root/reference/libs/datasets/loc_generators.py:61-79."""
    pts = []
    for l in range(n_levels):
        s = scale_factor ** l
        t = torch.arange(0, T, s, dtype=torch.float32)[:, None]
        rr = torch.tensor(regression_range[l], dtype=torch.float32)[None].repeat(t.shape[0], 1)
        st = torch.full((t.shape[0], 1), float(s))
        pts.append(torch.cat((t, rr, st), dim=1))
    return pts


def make_batch(B: int, T: int = 224, first_index: int = 0, num_classes: int = 100,
               n_levels: int = 6, len_lo: int = 60, len_hi: int = 187,
               with_gt: bool = True) -> dict:
    """Synthetic collate dict (CPU tensors).  GT tensors are filled with one plausible event per
    video so that the reference's unconditional loss code runs; the hot path ignores them."""
    len_hi = min(len_hi, T)
    vis = torch.zeros(B, 2048, T)
    aud = torch.zeros(B, 128, T)
    mask = torch.zeros(B, 1, T, dtype=torch.bool)
    lens = []
    for i in range(B):
        vid = first_index + i
        g = torch.Generator().manual_seed(1234 + vid)
        L = int(torch.randint(len_lo, len_hi + 1, (1,), generator=g).item())
        lens.append(L)
        vis[i, :, :L] = 0.3 * torch.randn(2048, L, generator=g).abs()
        aud[i, :, :L] = 0.5 * torch.randn(128, L, generator=g).abs()
        mask[i, 0, :L] = True
    Ttot = sum(T >> l for l in range(n_levels))
    batch = {
        "visual": vis, "audio": aud, "mask": mask,
        "video_id": [f"synth_{first_index + i:06d}" for i in range(B)],
        "fps": [25.0] * B,
        "duration": [(L * 8 + 24) / 25.0 for L in lens],
        "feat_stride": [8] * B,
        "feat_num_frames": [24] * B,
        "points": [p[None].repeat(B, 1, 1) for p in make_points(T, n_levels)],
        "lengths": lens,
    }
    if with_gt:
        scores = torch.zeros(B, T)
        start_end = torch.zeros(B, T)
        m_labels = torch.zeros(B, T, num_classes)
        for i, L in enumerate(lens):
            s, e = L // 4, L // 2
            scores[i, s:e] = 1
            start_end[i, s:e + 1] = 1
            m_labels[i, s:e, (first_index + i) % num_classes] = 1
        batch.update({
            "scores": scores, "start_end": start_end, "m_labels": m_labels,
            "gt_offsets": torch.zeros(B, Ttot, num_classes, 2),
            "gt_cls_labels": torch.zeros(B, Ttot, num_classes),
        })
    return batch


def add_event_targets(batch: dict, first_index: int = 0, key_labels=None, num_classes: int = 100,
                      regression_range=((0, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 10000))) -> dict:
    """Replace the placeholder GT tensors of ``make_batch`` with 1-3 seeded events per video, in the layout the reference's
    dataset + collate produce (/root/reference/libs/datasets/data_utils.py:141-160 for ``scores`` / ``start_end`` /
    ``m_labels``; per-point ``gt_cls_labels`` / ``gt_offsets`` = distances to the event boundaries in units of the level's
    stride, positive where the point lies inside the event and the larger distance falls in the level's regression range).
    Only the loss-only tail of the reference forward reads these tensors (multimodal_meta_archs.py:504-509).
    ``key_labels[i]``, if given, is the class of video i's first event."""
    lens = batch["lengths"]
    B, T = batch["mask"].shape[0], batch["mask"].shape[-1]
    pts = torch.cat([p[0] for p in batch["points"]], dim=0)                    # [Ttot, 4] = (t, lo, hi, stride)
    Ttot = pts.shape[0]
    scores = torch.zeros(B, T)
    start_end = torch.zeros(B, T)
    m_labels = torch.zeros(B, T, num_classes)
    gt_off = torch.zeros(B, Ttot, num_classes, 2)
    gt_cls = torch.zeros(B, Ttot, num_classes)
    for i, L in enumerate(lens):
        g = torch.Generator().manual_seed(99991 + first_index + i)
        n_ev = int(torch.randint(1, 4, (1,), generator=g))
        starts = sorted(float(x) for x in (2 + torch.rand(n_ev, generator=g) * (L - 16)))
        for e, s in enumerate(starts):
            dur = 4.0 + float(torch.rand(1, generator=g)) * (L / 2 - 4)
            end = min(s + dur, L - 1.0)
            c = int(torch.randint(0, num_classes, (1,), generator=g))
            if e == 0 and key_labels is not None:
                c = int(key_labels[i])
            si, ei = int(s), int(end)
            scores[i, si:ei] = 1
            start_end[i, si:ei + 1] = 1
            m_labels[i, si:ei] = 0
            m_labels[i, si:ei, c] = 1
            left = (pts[:, 0] - s) / pts[:, 3]
            right = (end - pts[:, 0]) / pts[:, 3]
            far = torch.maximum(left, right)
            pos = (left >= 0) & (right >= 0) & (far >= pts[:, 1]) & (far <= pts[:, 2]) & (pts[:, 0] < L)
            gt_cls[i, pos, c] = 1
            gt_off[i, pos, c, 0] = left[pos]
            gt_off[i, pos, c, 1] = right[pos]
    batch.update({"scores": scores, "start_end": start_end, "m_labels": m_labels, "gt_offsets": gt_off,
                  "gt_cls_labels": gt_cls})
    return batch


def make_items(B: int, first_index: int = 0, len_lo: int = 60, len_hi: int = 187) -> list:
    """The same synthetic videos as ``make_batch`` as a list of un-collated dataset items (ragged features), the
    input of the reference's ``collate_fcn`` (/root/reference/libs/datasets/data_utils.py:123) and of
    ``ingest.DeviceCollator``.  Inference keys only."""
    items = []
    for i in range(B):
        vid = first_index + i
        g = torch.Generator().manual_seed(1234 + vid)
        L = int(torch.randint(len_lo, len_hi + 1, (1,), generator=g).item())
        vis = 0.3 * torch.randn(2048, L, generator=g).abs()
        aud = 0.5 * torch.randn(128, L, generator=g).abs()
        items.append({"video_id": f"synth_{vid:06d}", "feats": {"visual": vis, "audio": aud}, "fps": 25.0,
                      "duration": (L * 8 + 24) / 25.0, "feat_stride": 8, "feat_num_frames": 24})
    return items
