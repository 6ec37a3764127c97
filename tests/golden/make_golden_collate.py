"""Golden fixture for the collate path: the REAL reference ``collate_fcn`` (libs/datasets/data_utils.py:123) on three
small ragged synthetic videos, eval branch.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_collate.py        -> tests/golden/collate_b3.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import import_reference  # noqa: E402


def main():
    import_reference()
    from libs.datasets.data_utils import collate_fcn
    g = torch.Generator().manual_seed(42)
    num_classes, T, n_levels = 5, 32, 6
    lens = [11, 32, 7]
    items, raw = [], {}
    Ttot = sum(T >> l for l in range(n_levels))
    for i, L in enumerate(lens):
        vis = torch.randn(16, L, generator=g)
        aud = torch.randn(4, L, generator=g)
        raw[f"visual_{i}"], raw[f"audio_{i}"] = vis.numpy(), aud.numpy()
        items.append({"video_id": f"v{i}", "feats": {"visual": vis, "audio": aud}, "fps": 25.0, "duration": L * 0.32 + 0.96,
                      "feat_stride": 8, "feat_num_frames": 24, "segments": torch.tensor([[1.0, 2.5]]),
                      "labels": torch.tensor([i % num_classes]), "gt_offsets": torch.zeros(Ttot, num_classes, 2),
                      "gt_cls_labels": torch.zeros(Ttot, num_classes), "points": [torch.zeros(T >> l, 4) for l in range(n_levels)]})
    out = collate_fcn(items, num_classes, T, padding_val=0, training=False, max_div_factor=1)
    np.savez_compressed(os.path.join(HERE, "collate_b3.npz"), lens=np.array(lens), max_seq_len=T, visual=out["visual"].numpy(),
                        audio=out["audio"].numpy(), mask=out["mask"].numpy(), **raw)
    print("wrote collate_b3.npz", out["visual"].shape, out["mask"].shape, out["mask"].sum(-1).flatten().tolist())


if __name__ == "__main__":
    main()
