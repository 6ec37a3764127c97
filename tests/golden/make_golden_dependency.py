"""Golden fixture for the Dependency_Block: the REAL reference module
(libs/modeling/dependency_block.py) with the name-keyed synthetic weights, two pyramid levels, one padded video.
Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_dependency.py       -> tests/golden/dependency_b2.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import import_reference  # noqa: E402
from unav_yolyolva_b200 import synth  # noqa: E402


def inputs():
    g = torch.Generator().manual_seed(77)
    T = 32
    feats = [torch.randn(2, 1024, T, generator=g), torch.randn(2, 1024, T // 2, generator=g)]
    m0 = torch.ones(2, 1, T, dtype=torch.bool)
    m0[1, 0, 21:] = False
    masks = [m0, m0[:, :, ::2]]
    feats = [f * m for f, m in zip(feats, masks)]
    return feats, masks


def main():
    import_reference()
    from libs.modeling import make_dependency_block
    torch.set_num_threads(os.cpu_count() or 1)
    blk = make_dependency_block("DependencyBlock", in_channel=1024, n_embd=128, n_embd_ks=3, num_classes=100, path_pdrop=0.1).eval()
    sd = {k: synth.trained_like_tensor("dependency_block." + k, list(v.shape)) for k, v in blk.state_dict().items()}
    blk.load_state_dict(sd, strict=True)
    feats, masks = inputs()
    with torch.no_grad():
        outs, _ = blk(feats, masks)
    np.savez_compressed(os.path.join(HERE, "dependency_b2.npz"), out0=outs[0].numpy(), out1=outs[1].numpy(),
                        names=np.array(sorted(sd)), shapes=np.array([str(list(sd[k].shape)) for k in sorted(sd)]))
    print("wrote dependency_b2.npz", [tuple(o.shape) for o in outs], [float(o.abs().max()) for o in outs],
          "NaN:", [bool(torch.isnan(o).any()) for o in outs])


if __name__ == "__main__":
    main()
