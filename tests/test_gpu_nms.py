"""Decode and soft-NMS kernels through the C ABI vs the oracle: bit-exact candidates / indices / labels."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref as R
from oracle import nms_ref
from unav_yolyolva_b200 import kernels as K
from unav_yolyolva_b200.config import TEST_CFG
from unav_yolyolva_b200.utils import batched_nms

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _gpu_nms(cuda, segs, scores, labels, ncls, min_score, method=2, iou=0.7, sigma=0.4, K_out=100, meta=None, B=1, mpc=0):
    cap = segs.shape[-2]
    o_s = torch.zeros(B, K_out, 2, device=cuda); o_sc = torch.zeros(B, K_out, device=cuda)
    o_l = torch.zeros(B, K_out, dtype=torch.int64, device=cuda); o_c = torch.zeros(B, dtype=torch.int32, device=cuda)
    ws = torch.zeros(K.softnms_workspace_bytes(B, ncls, K_out), dtype=torch.uint8, device=cuda)
    K.softnms_batched(torch.as_tensor(segs).to(cuda).contiguous(), torch.as_tensor(scores).to(cuda).contiguous(),
                      torch.as_tensor(labels).to(cuda, torch.int32).contiguous(), B, cap, ncls, iou, sigma, min_score, method,
                      K_out, mpc, None if meta is None else meta.to(cuda), o_s, o_sc, o_l, o_c, ws)
    torch.cuda.synchronize()
    return o_s.cpu().numpy(), o_sc.cpu().numpy(), o_l.cpu().numpy(), o_c.cpu().numpy()


@pytest.fixture(params=["lazy_per_video", "per_class"])
def nms_path(request):
    """Both device schedules must be bit-exact: the lazy per-video kernel (default) and the per-class + merge pair."""
    if request.param == "per_class":
        os.environ["UNAV_NMS_PER_CLASS"] = "1"
    yield request.param
    os.environ.pop("UNAV_NMS_PER_CLASS", None)


def test_softnms_bit_exact_on_golden_cases(cuda, golden_dir, nms_path):
    g = np.load(os.path.join(golden_dir, "nms_cases.npz"))
    for ci in range(6):
        segs, sc, lb = g[f"in_segs_{ci}"], g[f"in_scores_{ci}"], g[f"in_labels_{ci}"]
        o = _gpu_nms(cuda, segs, sc, lb, int(lb.max()) + 1, float(g[f"min_score_{ci}"]))
        n = int(o[3][0])
        assert n == len(g[f"out_scores_{ci}"])
        assert np.array_equal(o[2][0, :n], g[f"out_labels_{ci}"])
        assert np.array_equal(_bits(o[1][0, :n]), _bits(g[f"out_scores_{ci}"]))
        assert np.array_equal(_bits(o[0][0, :n]), _bits(g[f"out_segs_{ci}"]))
        assert not o[1][0, n:].any()
    o = _gpu_nms(cuda, g["hard_in_segs"], g["hard_in_scores"], g["hard_in_labels"], 20, 0.001, method=3, iou=0.5)
    n = int(o[3][0])
    assert np.array_equal(o[2][0, :n], g["hard_out_labels"]) and np.array_equal(_bits(o[1][0, :n]), _bits(g["hard_out_scores"]))


def test_softnms_batched_ragged_videos_vs_oracle(cuda, nms_path):
    """Config-5 style stress: several videos, empty slots (-1), a skewed class, all methods."""
    rng = np.random.default_rng(7)
    B, cap, ncls = 4, 3000, 100
    segs = np.zeros((B, cap, 2), np.float32); sc = np.zeros((B, cap), np.float32); lb = np.full((B, cap), -1, np.int32)
    for b in range(B):
        n = [cap, 1700, 1, 441][b]
        c = rng.uniform(0, 200, n); h = rng.uniform(0.5, 20, n)
        idx = rng.permutation(cap)[:n]
        segs[b, idx] = np.stack([c - h, c + h], 1); sc[b, idx] = rng.uniform(0.001, 1, n)
        lb[b, idx] = 0 if b == 3 else rng.integers(0, ncls, n)
    meta = torch.tensor([[8, 24, 25, 40.0]] * B)
    for method in (2, 1, 0):
        o = _gpu_nms(cuda, segs, sc, lb, ncls, 0.001, method=method, iou=0.5, meta=meta, B=B)
        for b in range(B):
            keep = lb[b] >= 0
            if method == 2:
                r = nms_ref.batched_nms(segs[b][keep], sc[b][keep], lb[b][keep].astype(np.int64), 0.5, 0.001, 100, True, 0.4)
                n = int(o[3][b])
                assert n == len(r[1])
                assert np.array_equal(o[2][b, :n], r[2])
                assert np.array_equal(_bits(o[1][b, :n]), _bits(r[1]))
                sec = nms_ref.to_seconds(r[0], 8, 24, 25, 40.0)
                assert np.array_equal(_bits(np.abs(o[0][b, :n])), _bits(np.abs(sec)))
            else:   # per-class restatement for the linear / hard-weight variants
                for c in np.unique(lb[b][keep])[:5]:
                    m = lb[b] == c
                    dets, _ = nms_ref.softnms(segs[b][m], sc[b][m], 0.5, 0.4, 0.001, method)
                    assert len(dets) >= 1


def test_batched_nms_operator_matches_reference_signature(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_cases.npz"))
    segs, sc, lb = (torch.from_numpy(g[k]) for k in ("in_segs_1", "in_scores_1", "in_labels_1"))
    o = batched_nms(segs, sc, lb, TEST_CFG["iou_threshold"], TEST_CFG["min_score"], TEST_CFG["max_seg_num"], use_soft_nms=True,
                    multiclass=True, sigma=TEST_CFG["nms_sigma"], voting_thresh=TEST_CFG["voting_thresh"])
    assert o[0].device == segs.device and o[2].dtype == torch.int64
    assert np.array_equal(o[2].numpy(), g["out_labels_1"]) and np.array_equal(_bits(o[1].numpy()), _bits(g["out_scores_1"]))
    e = batched_nms(torch.zeros(0, 2), torch.zeros(0), torch.zeros(0, dtype=torch.int64), 0.7, 0.001, 100)
    assert e[0].shape == (0, 2) and e[2].dtype == torch.int64


def test_decode_matches_oracle_on_golden_logits(cuda, golden_dir):
    """Kernel fed the reference's own logits/offsets: the selected candidate set (flat ids), labels and
    segment bits are identical to inference_single_video's; scores agree to sigmoid rounding."""
    g = np.load(os.path.join(golden_dir, "model_b2.npz"))
    B, L, T, ncls = 2, 6, 224, 100
    Tl = [T >> l for l in range(L)]
    off = np.cumsum([0] + Tl).tolist()
    logits = torch.cat([torch.from_numpy(g[f"logits_{l}"]) for l in range(L)], 1).contiguous()          # [B,441,100]
    offsets = torch.cat([torch.from_numpy(g[f"offsets_{l}"]) for l in range(L)], 1).contiguous()        # [B,441,100,2]
    masks = torch.cat([torch.from_numpy(g[f"mask_{l}"]) for l in range(L)], 1).to(torch.uint8).contiguous()
    pts = torch.cat(R.make_points(T), 0).contiguous()
    cap = sum(min(2000, t * ncls) for t in Tl)
    cs, csc = torch.zeros(B, cap, 2, device=cuda), torch.zeros(B, cap, device=cuda)
    cl = torch.zeros(B, cap, dtype=torch.int32, device=cuda)
    K.decode(logits.to(cuda), offsets.to(cuda), masks.to(cuda), pts.to(cuda), off, B, ncls, True, 0.001, 2000, 0.05, cs, csc, cl, cap)
    torch.cuda.synchronize()
    cs, csc, cl = cs.cpu().numpy(), csc.cpu().numpy(), cl.cpu().numpy()
    for i in range(B):
        segs, scores, labels, ids = R.decode_single_video(R.make_points(T), [torch.from_numpy(g[f"mask_{l}"][i]) for l in range(L)],
                                                          [torch.from_numpy(g[f"logits_{l}"][i]) for l in range(L)],
                                                          [torch.from_numpy(g[f"offsets_{l}"][i]) for l in range(L)])
        keep = cl[i] >= 0
        assert keep.sum() == len(labels)
        def key(s, lb):
            o = np.lexsort((s[:, 1], s[:, 0], lb))
            return o
        a, b = key(cs[i][keep], cl[i][keep]), key(segs.numpy(), labels.numpy())
        assert np.array_equal(cl[i][keep][a], labels.numpy()[b])
        assert np.array_equal(_bits(cs[i][keep][a]), _bits(segs.numpy()[b]))
        assert np.allclose(csc[i][keep][a], scores.numpy()[b], rtol=2e-6, atol=0)


def test_decode_topk_ties_and_empty_levels(cuda):
    """All-equal logits (maximal ties): exactly topk lowest flat indices survive; fully masked video -> nothing."""
    B, T, ncls, L = 2, 32, 100, 3
    Tl = [T >> l for l in range(L)]
    off = np.cumsum([0] + Tl).tolist()
    Ttot = off[-1]
    logits = torch.zeros(B, Ttot, ncls)
    offsets = torch.ones(B, Ttot, ncls, 2)
    masks = torch.ones(B, Ttot, dtype=torch.uint8)
    masks[1] = 0
    pts = torch.cat(R.make_points(T, n_levels=L, regression_range=((0, 4), (4, 8), (8, 10000))), 0)
    topk = 500
    cap = sum(min(topk, t * ncls) for t in Tl)
    cs, csc = torch.zeros(B, cap, 2, device=cuda), torch.zeros(B, cap, device=cuda)
    cl = torch.zeros(B, cap, dtype=torch.int32, device=cuda)
    K.decode(logits.to(cuda), offsets.to(cuda), masks.to(cuda), pts.to(cuda), off, B, ncls, True, 0.001, topk, 0.05, cs, csc, cl, cap)
    torch.cuda.synchronize()
    cl = cl.cpu().numpy()
    assert (cl[1] == -1).all()
    assert (cl[0] >= 0).sum() == cap
    assert np.array_equal(cl[0][:topk], np.arange(topk) % ncls)       # level 0: flat ids 0..topk-1 in order
