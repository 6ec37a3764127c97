"""Module-level forwards (reference calling convention, channels-first) vs the oracle restatement of the same
reference module, sharing the module's own parameters."""
import math

import pytest
import torch

from oracle import model_ref as R
from unav_yolyolva_b200 import _fwd
from unav_yolyolva_b200.modeling import blocks as BL
from unav_yolyolva_b200.modeling import multimodal_backbones as MB
from unav_yolyolva_b200.modeling import multimodal_meta_archs as MA

pytestmark = pytest.mark.gpu


def _rand_init(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if p.dim() >= 2 and p.shape[0] > 1 and not (p.dim() == 3 and p.shape[0] == 1):
                fan = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan))
            elif n.endswith("weight") or n.endswith("scale"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return mod


def _sd(mod, prefix="m"):
    return {f"{prefix}.{k}": v.detach().cpu() for k, v in mod.state_dict().items()}


def _inputs(B, C, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, T, generator=g)
    lens = torch.randint(T // 3, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] < lens[:, None])[:, None]
    return x, mask


def _close(a, b, tol):
    return float((a.cpu() - b).abs().max() / (b.abs().max() + 1e-12)) < tol


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-6), ("bf16x3", 5e-5)])
def test_masked_conv_variants(cuda, mode, tol):
    _fwd.MODE = mode
    x, mask = _inputs(3, 256, 56)
    for kw in (dict(kernel_size=3, padding=1), dict(kernel_size=1), dict(kernel_size=3, padding=1, stride=2),
               dict(kernel_size=3, padding=1, groups=256, bias=False), dict(kernel_size=3, padding=1, groups=256, bias=False, stride=2)):
        cout = 256 if kw.get("groups") else 200
        m = _rand_init(BL.MaskedConv1D(256, cout, **kw), 1).to(cuda)
        y, mo = m(x.to(cuda), mask.to(cuda))
        ry, rm = R.masked_conv1d(_sd(m), "m", x, mask, stride=kw.get("stride", 1), groups=kw.get("groups", 1))
        assert _close(y, ry, tol) and torch.equal(mo.cpu(), rm)
    _fwd.MODE = "bf16x3"


def test_layernorm_and_mhca_and_block(cuda):
    _fwd.MODE = "fp32"
    for C, nh, T in ((256, 4, 112), (512, 4, 224)):
        x, mask = _inputs(2, C, T, seed=C)
        x2, _ = _inputs(2, C, T, seed=C + 1)
        ln = _rand_init(BL.LayerNorm(C), 2).to(cuda)
        assert _close(ln(x.to(cuda)), R.channel_ln(_sd(ln), "m", x), 2e-6)
        att = _rand_init(BL.MaskedMHCA(C, nh), 3).to(cuda)
        y, _ = att(x.to(cuda), x2.to(cuda), mask.to(cuda))
        ry, _ = R.masked_mhca(_sd(att), "m", x, x2, mask, nh)
        assert _close(y, ry, 5e-6)
        tb = _rand_init(BL.TransformerBlock(C, nh, path_pdrop=0.1), 4).to(cuda).eval()
        xd = x.to(cuda)
        y, _ = tb(xd, xd, mask.to(cuda))
        ry, _ = R.transformer_block(_sd(tb), "m", x, x, mask, nh)
        assert _close(y, ry, 5e-6)
    _fwd.MODE = "bf16x3"


def test_pyramid_fusion_pieces(cuda):
    _fwd.MODE = "fp32"
    x, mask = _inputs(2, 512, 56, seed=7)
    d = _rand_init(MB.Downsample_pyramid_levels(512, 2), 5).to(cuda)
    y, mo = d(x.to(cuda), mask.to(cuda))
    ry, rm = R.masked_conv1d(_sd(d), "m.down_conv", x, mask, stride=2, groups=512)
    ry = R.channel_ln(_sd(d), "m.down_norm", ry)
    assert _close(y, ry, 3e-6) and torch.equal(mo.cpu(), rm)
    ds = _rand_init(MB.downsample(512, 2), 6).to(cuda)
    y, _ = ds(x.to(cuda), mask.to(cuda))
    ry, _ = R.masked_conv1d(_sd(ds), "m.down_conv", x, mask, stride=2)
    ry = torch.nn.functional.silu(R.channel_ln(_sd(ds), "m.down_norm", ry))
    assert _close(y, ry, 5e-6)
    guide = torch.randn(2, 512, 224)
    xa, maska = _inputs(2, 256, 56, seed=9)
    for H in (8, 4):
        ab = _rand_init(MB.MaxSigmoidAttnBlock(256, 256, guide_channels=224, embed_channels=256, num_heads=H), 7).to(cuda)
        y, _ = ab(xa.to(cuda), guide.to(cuda), maska.to(cuda))
        ry, _ = R.maxsig_attn_block(_sd(ab), "m", xa, guide, maska, H)
        assert _close(y, ry, 5e-6)
    csp = _rand_init(MB.MaxSigmoidCSPLayerWithTwoConv(1024, 512, guide_channels=224, embed_channels=256, num_heads=8, num_blocks=3), 8).to(cuda)
    xc, maskc = _inputs(2, 1024, 28, seed=11)
    y, _ = csp(xc.to(cuda), guide.to(cuda), maskc.to(cuda))
    ry, _ = R.csp_layer(_sd(csp), "m", xc, guide, maskc, 8)
    assert _close(y, ry, 1e-5)
    _fwd.MODE = "bf16x3"


def test_heads_and_inference_entry(cuda, golden_dir):
    import os
    import numpy as np
    from unav_yolyolva_b200 import synth
    from unav_yolyolva_b200.config import default_model_cfg
    from unav_yolyolva_b200.modeling import make_multimodal_meta_arch
    _fwd.MODE = "fp32"
    model = make_multimodal_meta_arch("LocPointTransformer", **default_model_cfg())
    sd = synth.trained_like_state_dict()
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    B = 2
    feats, masks = [], []
    for l, T in enumerate((28, 14)):
        x, m = _inputs(B, 1024, T, seed=20 + l)
        feats.append(x); masks.append(m)
    lg = model.cls_head([f.to(cuda) for f in feats], [m.to(cuda) for m in masks])
    of = model.reg_head([f.to(cuda) for f in feats], [m.to(cuda) for m in masks])
    for l in range(2):
        c = r = feats[l]
        for i in range(2):
            c, _ = R.masked_conv1d(sd, f"cls_head.head.{i}", c, masks[l]); c = torch.relu(R.channel_ln(sd, f"cls_head.norm.{i}", c))
            r, _ = R.masked_conv1d(sd, f"reg_head.head.{i}", r, masks[l]); r = torch.relu(R.channel_ln(sd, f"reg_head.norm.{i}", r))
        c, _ = R.masked_conv1d(sd, "cls_head.cls_head", c, masks[l])
        r, _ = R.masked_conv1d(sd, "reg_head.offset_head", r, masks[l])
        r = torch.relu(r * sd[f"reg_head.scale.{l}.scale"])
        assert _close(lg[l], c, 5e-6) and _close(of[l], r, 5e-6)
    # inference() entry fed the reference's own golden head outputs -> the reference's golden detections
    g = np.load(os.path.join(golden_dir, "model_b2.npz"))
    batch = synth.make_batch(2, 224)
    res = model.inference(batch, [torch.from_numpy(g[f"mask_{l}"]).to(cuda) for l in range(6)],
                          [torch.from_numpy(g[f"logits_{l}"]).to(cuda) for l in range(6)],
                          [torch.from_numpy(g[f"offsets_{l}"]).to(cuda) for l in range(6)])
    assert np.array_equal(res["labels"].cpu().numpy(), g["labels"])
    assert np.allclose(res["scores"].cpu().numpy(), g["scores"], rtol=2e-5, atol=1e-7)
    assert np.allclose(res["segments"].cpu().numpy(), g["segments"], atol=2e-3)
    _fwd.MODE = "bf16x3"


def test_alignment_and_backbone_modules(cuda):
    from unav_yolyolva_b200 import synth
    _fwd.MODE = "fp32"
    sd = synth.trained_like_state_dict()
    b = synth.make_batch(2, 224)
    al = MB.Alignment(video_dim=2048, audio_dim=128)
    al.load_state_dict({k[len("alignment."):]: v for k, v in sd.items() if k.startswith("alignment.")}, strict=True)
    al = al.to(cuda).eval()
    v, a, _ = al(video=[b["visual"].to(cuda)], text=[b["audio"].to(cuda)], mask_video=[b["mask"].to(cuda)],
                 mask_text=[b["mask"].to(cuda)], m_start_end=None, m_scores_gt=None, m_labels=None)
    rv, ra = R.alignment(sd, "alignment", b["visual"], b["audio"], b["mask"])
    assert _close(v[0], rv, 1e-5) and _close(a[0], ra, 1e-5)
    bb = MB.ConvTransformerBackbone(512, 512, 512, 4, 3, 224, arch=(2, 3, 5), scale_factor=2, with_ln=True, path_pdrop=0.1,
                                    use_abs_pe=True)
    bb.load_state_dict({k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}, strict=True)
    bb = bb.to(cuda).eval()
    fv, fa, ms = bb(rv.to(cuda), ra.to(cuda), b["mask"].to(cuda))
    sd2 = dict(sd); sd2["backbone.pos_embd"] = R.sinusoid_pos_embd(224, 512)
    rfv, rfa, rms = R.backbone(sd2, "backbone", rv, ra, b["mask"])
    for l in range(6):
        assert _close(fv[l], rfv[l], 2e-5) and _close(fa[l], rfa[l], 2e-5)
        assert torch.equal(ms[l].cpu(), rms[l])
    _fwd.MODE = "bf16x3"


def test_dependency_block_matches_reference_golden(cuda, golden_dir):
    """Dependency_Block through the module-level kernels vs the reference module's own output:
    two pyramid levels, one padded video, the per-sequence-mask co-occurrence branch, the tiled temporal mask."""
    import os
    import numpy as np
    from unav_yolyolva_b200 import synth
    from unav_yolyolva_b200.modeling import make_dependency_block
    d = np.load(os.path.join(golden_dir, "dependency_b2.npz"))
    blk = make_dependency_block("DependencyBlock", in_channel=1024, n_embd=128, n_embd_ks=3, num_classes=100, path_pdrop=0.1)
    sd = {k: synth.trained_like_tensor("dependency_block." + k, list(v.shape)) for k, v in blk.state_dict().items()}
    assert sorted(sd) == [str(n) for n in d["names"]]
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(cuda).eval()
    g = torch.Generator().manual_seed(77)
    T = 32
    feats = [torch.randn(2, 1024, T, generator=g), torch.randn(2, 1024, T // 2, generator=g)]
    m0 = torch.ones(2, 1, T, dtype=torch.bool)
    m0[1, 0, 21:] = False
    masks = [m0, m0[:, :, ::2]]
    feats = [f * m for f, m in zip(feats, masks)]
    with torch.no_grad():
        outs, _ = blk([f.to(cuda) for f in feats], [m.to(cuda) for m in masks])
    torch.cuda.synchronize()
    for o, key in zip(outs, ("out0", "out1")):
        ref = torch.from_numpy(d[key])
        err = float((o.cpu() - ref).abs().max() / ref.abs().max())
        print(f"dependency block {key}: rel err {err:.3e}")
        assert err < 2e-4
