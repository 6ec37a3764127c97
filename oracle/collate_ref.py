"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the inference part of the reference's ``collate_fcn``
(/root/reference/libs/datasets/data_utils.py:123-229): padding of the per-video ``[C, len]`` features to a common
length (:170-198) and the validity mask ``arange(max_len) < len`` (:201-203).  Pinned against the reference's own
``collate_fcn`` run in the build container (tests/golden/collate_b3.npz, tests/golden/make_golden.py).
Only tests/ may import this module."""
import numpy as np


def collate_pad(feats, max_seq_len, padding_val=0.0, max_div_factor=1):
    """feats: list of [C, len_i] float32 arrays -> (padded [B, C, T] float32, mask [B, 1, T] bool)."""
    lens = np.array([f.shape[-1] for f in feats])
    max_len = int(lens.max())
    if max_len <= max_seq_len:                       # eval branch, :170-176
        T = max_seq_len
    else:
        T = (max_len + (max_div_factor - 1)) // max_div_factor * max_div_factor
    out = np.full((len(feats), feats[0].shape[0], T), padding_val, dtype=np.float32)
    for f, o in zip(feats, out):                     # :181-182, :197-198
        o[..., :f.shape[-1]] = f
    mask = (np.arange(T)[None, :] < lens[:, None])[:, None, :]     # :201-203
    return out, mask
