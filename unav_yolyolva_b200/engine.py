"""Fused, token-major execution of the PtTransformer inference hot path on sm_100a.

``HotPathEngine`` turns a ``PtTransformer`` (parameter holder with the reference's state_dict layout) into
a static launch plan over hand-written CUDA kernels (C ABI, ``kernels.py``): Alignment multiway transformer
-> conv/transformer stem -> depthwise pyramid -> two fusion passes (batched as 2B) -> cls/reg heads -> decode
-> per-class soft-NMS.  256 kernel launches per batch replace the reference's ~51 k ATen calls
(SURVEY.md §2.2); the whole sequence is captured once per batch size into a CUDA graph and replayed.

Data layout: activations are token-major FP32 ``[rows, C]`` matrices (rows = time steps of
consecutive batch items); the two modalities are stacked along the batch ("NB = 2B": visual items first), the
6 pyramid levels are separate matrices until the heads, where they are concatenated per video (441 rows).
GEMM operands are separate buffers in the operand dtype of the chosen precision mode:

    mode      operands              GEMM backend                    max |err| / range on logits vs FP32 oracle
    fp32      FP32                  CUDA-core FFMA (gemm_simt)      ~1e-6
    f16x3     FP16 hi/lo split      tcgen05, 3 MMA passes           ~1e-6   (2 x 11 significand bits)
    bf16x3    BF16 hi/lo split      tcgen05, 3 MMA passes           ~8e-6   (2 x 8 bits, FP32 exponent range)
    fast      FP16 hi/lo split      tcgen05, 1 pass (hi.hi) in the Alignment and backbone stages,
                                    3 passes in the fusion passes and heads              ~2e-4 (north_star's BF16 budget: 1e-3)
    f16       FP16                  tcgen05, 1 MMA pass             ~4e-4 (offsets 1.2e-3)
    bf16      BF16                  tcgen05, 1 MMA pass             ~3e-3
The per-stage pass counts come from scripts/precision_study.py (oracle with emulated operand rounding).

k=3 convolutions are GEMMs over im2col operands that the producing kernel scatters directly (no separate
im2col pass for LayerNorm outputs).  Reference semantics restated per step with file:line in the comments
(paths relative to /root/reference/libs/modeling).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

from . import _cabi
from . import kernels as K
from .kernels import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SILU, BF16, BF16X2, F16, F16X2, F32, GEMM_SIMT, GEMM_TCGEN05, View

MODES = {"fp32": (F32, GEMM_SIMT), "bf16": (BF16, GEMM_TCGEN05), "bf16x3": (BF16X2, GEMM_TCGEN05),
         "f16": (F16, GEMM_TCGEN05), "f16x3": (F16X2, GEMM_TCGEN05), "fast": (F16X2, GEMM_TCGEN05),
         "bf16_simt": (BF16, GEMM_SIMT), "bf16x3_simt": (BF16X2, GEMM_SIMT)}
# MMA passes over split operands per stage of the path (0 = all three); stages not listed run all passes
STAGE_PASSES = {"fast": {"alignment": 1, "backbone": 1}}


class _Arena:
    """Zero-initialised device memory for a plan's buffers, carved from a few large chunks (one fill per chunk instead of
    one eager kernel per buffer); every buffer starts on a 256-byte boundary."""

    CHUNK = 64 << 20

    def __init__(self, device):
        self.dev, self.chunks, self.off = device, [], 0

    def alloc(self, shape, dtype) -> torch.Tensor:
        shape = tuple(int(x) for x in shape)
        n = 1
        for x in shape:
            n *= x
        es = torch.empty((), dtype=dtype).element_size()
        nbytes = (n * es + 255) // 256 * 256
        if not self.chunks or self.off + nbytes > self.chunks[-1].numel():
            self.chunks.append(torch.zeros(max(self.CHUNK, nbytes), dtype=torch.uint8, device=self.dev))
            self.off = 0
        t = self.chunks[-1][self.off:self.off + n * es].view(dtype).view(shape)
        self.off += nbytes
        return t


class HotPathEngine:
    def __init__(self, model, mode: str = "bf16x3", use_graph: bool = True):
        if mode not in MODES:
            raise ValueError(f"unknown precision mode {mode!r}; choose from {sorted(MODES)}")
        self.mode = mode
        self.op, self.backend = MODES[mode]
        # the Dependency_Block path (module-level kernels + torch views) is captured with the rest (its temporaries live in
        # the graph's private pool; bit-identical to eager, 29.4 -> 26.6 ms per batch of 16); UNAV_DEP_GRAPH=0 keeps it eager
        self.use_graph = use_graph and (not getattr(model, "use_dependency", False) or os.environ.get("UNAV_DEP_GRAPH", "1") != "0")
        self.model = model
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("HotPathEngine needs the model on a CUDA device (no CPU fallback)")
        self.dev = dev
        from . import _cabi
        _cabi.check(_cabi.load().unav_check_device(dev.index if dev.index is not None else torch.cuda.current_device()),
                    "unav_check_device")
        self._validate_structure(model)
        self.T = model.max_seq_len
        self.C = model.backbone.n_embd
        self.L = len(model.fpn_strides)
        self.ncls = model.num_classes
        self.n_head = model.backbone.n_head
        self.Tl = [self.T // s for s in model.fpn_strides]
        self.Ttot = sum(self.Tl)
        self.level_off = [0]
        for t in self.Tl:
            self.level_off.append(self.level_off[-1] + t)
        # fused tcgen05/TMEM attention for the tensor-core modes: one tile of keys up to 256 (every level of the T = 224 path),
        # key-chunked + merge above that (config 4, T = 2304); the CUDA-core kernel serves the fp32 / *_simt modes
        self.tc_attn = self.backend == GEMM_TCGEN05 and not (self.T + 1 > 256 and os.environ.get("UNAV_ATTN_LONG_SIMT") == "1")   # A/B knob
        self.attn_simt_max_t = int(os.environ.get("UNAV_ATTN_SIMT_MAX_T", "0"))      # A/B knob, see _tc_attn_for
        self.implicit_c3 = os.environ.get("UNAV_IMPLICIT_C3", "0") == "1"            # A/B knob, see _csp
        self.stage_passes = STAGE_PASSES.get(mode, {})
        self._stage = "alignment"
        self.w: Dict[str, torch.Tensor] = {}
        self._pack_weights()
        self._plans: Dict[tuple, dict] = {}

    @staticmethod
    def _validate_structure(model):
        """The launch plan hard-codes the structure of the reference configs (configs/avel_unav100.yaml over
        libs/core/config.py DEFAULTS).  Any other valid reference config would pack weights / launch GEMMs with the wrong K
        or layout and give silently wrong output, so it is rejected here with the offending field named."""
        from torch import nn
        bb = model.backbone

        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"HotPathEngine: unsupported model structure — {what} (the fused plan covers the "
                                          "reference's avel_unav100 configuration family only)")
        need(len(bb.embd_V) == 2 and len(bb.embd_A) == 2, f"backbone_arch[0] = {len(bb.embd_V)} embedding convs (need 2)")
        need(all(m.conv.kernel_size[0] == 3 and m.conv.bias is None for m in list(bb.embd_V) + list(bb.embd_A)),
             "embedding convs must be k = 3 without bias (embd_kernel_size = 3, embd_with_ln = True)")
        need(not isinstance(bb.embd_norm_V[0], nn.Identity), "embd_with_ln = False")
        need(bb.use_abs_pe, "use_abs_pe = False")
        need(bb.scale_factor == 2 and list(model.fpn_strides) == [2 ** i for i in range(len(model.fpn_strides))],
             f"scale_factor = {bb.scale_factor} (pyramid kernels are stride 2)")
        need(len(model.fpn_strides) >= 4, "fewer than 4 pyramid levels (guide enhancement pools the 3 finest)")
        need(bb.n_embd == 512 and bb.n_embd % bb.n_head == 0 and bb.n_embd // bb.n_head in (64, 128), f"embd_dim = {bb.n_embd}, n_head = {bb.n_head}")
        for h in (model.cls_head, model.reg_head):
            need(len(h.head) == 2, f"head_num_layers = {len(h.head) + 1} (need 3)")
            need(all(m.conv.kernel_size[0] == 3 and m.conv.bias is None for m in h.head), "head convs must be k = 3 with LayerNorm "
                 "(head_kernel_size = 3, head_with_ln = True)")
            need(not isinstance(h.norm[0], nn.Identity), "head_with_ln = False")
        need(model.cls_head.cls_head.conv.kernel_size[0] == 3 and model.reg_head.offset_head.conv.kernel_size[0] == 3, "head_kernel_size != 3")
        need(model.class_aware, "class_aware = False")
        need(model.alignment.num_layers >= 1 and model.alignment.multiway_list[0].attn_fusion._heads == 8, "Alignment with other than 8 heads")

    # ------------------------------------------------------------------------------ weights
    def _pack_weights(self):
        """Weights are prepared on the HOST (reshape / permute / concatenate of the state_dict's FP32 tensors), uploaded as one
        FP32 staging buffer, and converted into GEMM operand rows by ``unav_pack_operand`` (one launch per tensor) — no eager
        PyTorch kernel runs on the device at load time.  Vectors (biases, LayerNorm affine, scales) stay FP32 views of the
        uploaded buffer."""
        src = self.model.state_dict()

        class _HostSD(dict):                 # state_dict tensors as FP32 host tensors, fetched on first use
            def __missing__(d, k):
                d[k] = src[k].detach().to("cpu", torch.float32)
                return d[k]

        sd = _HostSD()
        w = self.w
        self._ops, self._vecs = [], []       # (name, host tensor [N, K]) / (name, host tensor any shape)

        def op_(name, t2d):
            self._ops.append((name, t2d.contiguous()))

        def lin(name, key):            # nn.Linear / Conv1d k=1: [N, K(,1)]
            t = sd[key]
            op_(name, t.reshape(t.shape[0], -1))

        def conv3(name, key):          # Conv1d k=3 [N, Cin, 3] -> im2col layout [N, 3*Cin] (tap-major)
            t = sd[key]
            op_(name, t.permute(0, 2, 1).reshape(t.shape[0], -1))

        def vec(name, key):
            self._vecs.append((name, sd[key].reshape(-1)))

        def cat_lin(name, keys):
            op_(name, torch.cat([sd[k].reshape(sd[k].shape[0], -1) for k in keys], 0))

        def cat_vec(name, keys):
            self._vecs.append((name, torch.cat([sd[k].reshape(-1) for k in keys])))

        self._sd, self._op_, self._vec_ = sd, op_, vec
        # ---- Alignment (multimodal_backbones.py:989-1034)
        a = "alignment."
        lin("al.pv", a + "proj_fc_video.0.weight"); vec("al.pv.b", a + "proj_fc_video.0.bias")
        lin("al.pa", a + "proj_fc_text.0.weight"); vec("al.pa.b", a + "proj_fc_text.0.bias")
        N = self.T + 1
        self._vecs.append(("al.pos_v", sd[a + "pos_embed_video"][0, :N].contiguous()))
        self._vecs.append(("al.pos_a", sd[a + "pos_embed_text"][0, :N].contiguous()))
        for nm in ("type_video", "type_text", "cls_token_video", "cls_token_text"):
            vec("al." + nm, a + nm)
        m = a + "multiway_list.0."
        vec("al.n1.w", m + "norm1_fused.weight"); vec("al.n1.b", m + "norm1_fused.bias")
        cat_lin("al.qkv", [m + f"attn_fusion.{x}.weight" for x in "qkv"])
        cat_vec("al.qkv.b", [m + f"attn_fusion.{x}.bias" for x in "qkv"])
        lin("al.m", m + "attn_fusion.m.weight"); vec("al.m.b", m + "attn_fusion.m.bias")
        for mod in ("video", "text"):
            vec(f"al.n2.{mod}.w", m + f"norm2_{mod}.weight"); vec(f"al.n2.{mod}.b", m + f"norm2_{mod}.bias")
            lin(f"al.fc1.{mod}", m + f"ffn_{mod}.fc1.weight"); vec(f"al.fc1.{mod}.b", m + f"ffn_{mod}.fc1.bias")
            lin(f"al.fc2.{mod}", m + f"ffn_{mod}.fc2.weight"); vec(f"al.fc2.{mod}.b", m + f"ffn_{mod}.fc2.bias")
            vec(f"al.nf.{mod}.w", a + f"norm_{mod}.weight"); vec(f"al.nf.{mod}.b", a + f"norm_{mod}.bias")
            lin(f"al.fc.{mod}", a + f"fc_{mod}.0.weight"); vec(f"al.fc.{mod}.b", a + f"fc_{mod}.0.bias")
            vec(f"al.fcn.{mod}.w", a + f"fc_{mod}.3.weight"); vec(f"al.fcn.{mod}.b", a + f"fc_{mod}.3.bias")

        # ---- backbone stem (multimodal_backbones.py:660-713)
        b = "backbone."
        for X in "VA":
            for i in range(2):
                conv3(f"bb.embd{X}{i}", b + f"embd_{X}.{i}.conv.weight")
                vec(f"bb.embdn{X}{i}.w", b + f"embd_norm_{X}.{i}.weight"); vec(f"bb.embdn{X}{i}.b", b + f"embd_norm_{X}.{i}.bias")
            for i in range(len(self.model.backbone.self_att_V)):
                self._pack_tblock(sd, f"bb.sa{X}{i}", b + f"self_att_{X}.{i}.")
        self._vecs.append(("bb.pe", self.model.backbone.pos_embd[0].detach().to("cpu", torch.float32).t().contiguous()))   # [T, C]
        for i in range(self.L - 1):
            vec(f"bb.down{i}.dw", b + f"downsample_list.{i}.down_conv.conv.weight")
            vec(f"bb.down{i}.w", b + f"downsample_list.{i}.down_norm.weight"); vec(f"bb.down{i}.b", b + f"downsample_list.{i}.down_norm.bias")
        # ---- fusion module (multimodal_backbones.py:367-549)
        f = b + "fusion_module."
        self._pack_mhca(sd, "fu.te", f + "text_enhancer.")
        conv3("fu.ds", f + "downsample_layers.0.down_conv.conv.weight"); vec("fu.ds.b", f + "downsample_layers.0.down_conv.conv.bias")
        vec("fu.dsn.w", f + "downsample_layers.0.down_norm.weight"); vec("fu.dsn.b", f + "downsample_layers.0.down_norm.bias")
        self._vecs.append(("fu.match.w", sd[f + "match_projection.weight"].reshape(sd[f + "match_projection.weight"].shape[0], -1).contiguous()))
        vec("fu.match.b", f + "match_projection.bias")
        self.td_heads, self.bu_heads = [], []
        for kind, heads in (("top_down_layers", self.td_heads), ("bottom_up_layers", self.bu_heads)):
            tag = "td" if kind.startswith("top") else "bu"
            for i in range(self.L - 1):
                p = f + f"{kind}.{i}."
                lin(f"fu.{tag}{i}.main", p + "main_conv.conv.weight"); vec(f"fu.{tag}{i}.main.b", p + "main_conv.conv.bias")
                for j in range(3):
                    self._pack_mhca(sd, f"fu.{tag}{i}.blk{j}", p + f"blocks.{j}.")
                vec(f"fu.{tag}{i}.hb", p + "attn_block.bias")
                heads.append(sd[p + "attn_block.bias"].numel())
                conv3(f"fu.{tag}{i}.proj", p + "attn_block.project_conv.conv.weight"); vec(f"fu.{tag}{i}.proj.b", p + "attn_block.project_conv.conv.bias")
                lin(f"fu.{tag}{i}.final", p + "final_conv.conv.weight"); vec(f"fu.{tag}{i}.final.b", p + "final_conv.conv.bias")
            cat_lin(f"fu.{tag}.gfc", [f + f"{kind}.{i}.attn_block.guide_fc.weight" for i in range(self.L - 1)])
            cat_vec(f"fu.{tag}.gfc.b", [f + f"{kind}.{i}.attn_block.guide_fc.bias" for i in range(self.L - 1)])
        # ---- heads (multimodal_meta_archs.py:101-259): first convs of cls and reg fused along N
        t0 = [sd[f"{h}.head.0.conv.weight"].permute(0, 2, 1).reshape(self.C, -1) for h in ("cls_head", "reg_head")]
        op_("hd.c0", torch.cat(t0, 0))
        for h, tag in (("cls_head", "cls"), ("reg_head", "reg")):
            conv3(f"hd.{tag}.c1", f"{h}.head.1.conv.weight")
            for i in range(2):
                vec(f"hd.{tag}.n{i}.w", f"{h}.norm.{i}.weight"); vec(f"hd.{tag}.n{i}.b", f"{h}.norm.{i}.bias")
        conv3("hd.cls.out", "cls_head.cls_head.conv.weight"); vec("hd.cls.out.b", "cls_head.cls_head.conv.bias")
        conv3("hd.reg.out", "reg_head.offset_head.conv.weight"); vec("hd.reg.out.b", "reg_head.offset_head.conv.bias")
        self.head_scales = [float(sd[f"reg_head.scale.{l}.scale"].reshape(-1)[0]) for l in range(self.L)]
        self._upload_weights()

    def _upload_weights(self):
        """One host -> device copy of everything registered by _pack_weights, then one pack kernel per operand."""
        dev, op, w = self.dev, self.op, self.w
        al = lambda n: (n + 63) // 64 * 64                   # 256-byte granules (FP32 elements)
        n_vec = sum(al(t.numel()) for _, t in self._vecs)
        n_src = sum(al(t.numel()) for _, t in self._ops)
        host = torch.empty(n_vec + n_src, dtype=torch.float32)
        o, where = 0, {}
        for name, t in self._vecs + self._ops:
            host[o:o + t.numel()].copy_(t.reshape(-1))
            where[name] = o
            o += al(t.numel())
        stage = host.to(dev)                                  # ONE memcpy
        self._w_vec = stage                                   # the vectors live in its first n_vec elements
        for name, t in self._vecs:
            w[name] = stage[where[name]:where[name] + t.numel()].view(t.shape)
        es = 4 if op == F32 else 2
        alo = lambda n: (n * es + 255) // 256 * 256 // es
        total = sum(alo(t.shape[0] * K.op_cols(t.shape[1], op)) for _, t in self._ops)
        arena = torch.empty(total, dtype=K.OP_TORCH_DTYPE[op], device=dev)        # fully written by the pack kernels
        self._w_op = arena
        o = 0
        for name, t in self._ops:
            N, Kc = t.shape
            cols = K.op_cols(Kc, op)
            dst = arena[o:o + N * cols].view(N, cols)
            K.pack_operand_into(stage[where[name]:where[name] + N * Kc].view(N, Kc), dst, op)
            w[name] = dst
            o += alo(N * cols)
        torch.cuda.current_stream(dev).synchronize()          # the FP32 staging copy of the operands may be reused afterwards
        self._ops = self._vecs = self._sd = self._op_ = self._vec_ = None

    def _pack_mhca(self, sd, name, p):
        for x in ("query", "key", "value"):
            self._vec_(f"{name}.{x}.dw", p + f"{x}_conv.conv.weight")
            self._vec_(f"{name}.{x}.nw", p + f"{x}_norm.weight"); self._vec_(f"{name}.{x}.nb", p + f"{x}_norm.bias")
            t = sd[p + f"{x}.weight"]
            self._op_(f"{name}.{x}", t.reshape(t.shape[0], -1))
            self._vec_(f"{name}.{x}.b", p + f"{x}.bias")
        t = sd[p + "proj.weight"]
        self._op_(f"{name}.proj", t.reshape(t.shape[0], -1))
        self._vec_(f"{name}.proj.b", p + "proj.bias")

    def _pack_tblock(self, sd, name, p):
        self._pack_mhca(sd, name + ".attn", p + "attn.")
        for ln in ("ln11", "ln12", "ln2"):
            self._vec_(f"{name}.{ln}.w", p + ln + ".weight"); self._vec_(f"{name}.{ln}.b", p + ln + ".bias")
        for i, tag in ((0, "mlp0"), (3, "mlp3")):
            t = sd[p + f"mlp.{i}.weight"]
            self._op_(f"{name}.{tag}", t.reshape(t.shape[0], -1))
            self._vec_(f"{name}.{tag}.b", p + f"mlp.{i}.bias")
        self._vec_(f"{name}.sa", p + "drop_path_attn.scale")
        self._vec_(f"{name}.sm", p + "drop_path_mlp.scale")

    # ------------------------------------------------------------------------------ buffers
    def _plan(self, B: int, slot: int = 0) -> dict:
        """Static buffers (and later the CUDA graphs, streams, events) of one batch size.  ``slot`` > 0: an independent
        second set, so that two batches can be in flight at once (streaming mode)."""
        if (B, slot) in self._plans:
            return self._plans[(B, slot)]
        dev, op, T, C, L = self.dev, self.op, self.T, self.C, self.L
        NB = 2 * B
        ar = _Arena(dev)
        f32 = lambda *s: ar.alloc(s, torch.float32)
        u8 = lambda *s: ar.alloc(s, torch.uint8)
        opb = lambda rows, Kc: ar.alloc((rows, K.op_cols(Kc, op)), K.OP_TORCH_DTYPE[op])
        P: dict = {"B": B, "arena": ar}
        Tl, Ttot = self.Tl, self.Ttot
        # inputs (static addresses for graph replay)
        P["visual"] = f32(B, 2048, T); P["audio"] = f32(B, 128, T); P["mask_in"] = u8(B, T)
        P["vid_meta"] = f32(B, 4)
        # masks
        P["m_true"] = u8(NB * Ttot); P["m_up"] = u8(NB * (Ttot - Tl[-1])); P["m_cls"] = u8(B, T + 1)
        offs, o = [], 0
        for l in range(L):
            offs.append(o); o += NB * Tl[l]
        P["m_off"] = offs
        # edge flags
        def edges(nseg, seglens):
            e = []
            for _ in range(nseg):
                for sl in seglens:
                    row = [0] * sl
                    row[0] |= 1; row[-1] |= 2
                    e.extend(row)
            return torch.tensor(e, dtype=torch.uint8, device=dev)
        P["edge_T"] = edges(NB, [T])
        P["edge_heads"] = edges(B, Tl)
        # alignment
        M_al = NB * (T + 1)
        P["Xv"] = opb(B * T, 2048); P["Xa"] = opb(B * T, 128)
        P["x0"] = f32(NB * T, C)
        P["F"] = f32(M_al, C); P["F1"] = f32(M_al, C)
        P["Fn"] = opb(M_al, C)
        P["QKV"] = f32(M_al, 3 * C)
        if self.tc_attn and T + 1 > 256:      # partial outputs + (max, sum) of the key chunks, largest call
            calls = [(2, B, T + 1, 8, C // 8), (1, NB, T, self.n_head, C // self.n_head), (1, NB, T, 4, C // 4), (1, NB, T, 4, C // 8)]
            P["att_ws"] = u8(max(K.attention_tc_workspace_bytes(ng, nb_, t_, t_, nh_, hs_) for ng, nb_, t_, nh_, hs_ in calls))
        if self.tc_attn:
            P["QKVop"] = opb(M_al, 3 * C)
            P["VTa"] = opb(NB * C, T + 1)          # values of both modalities transposed per item: [2B*C, T+1]
        P["AOa"] = opb(M_al, C)
        P["Ha"] = opb(M_al, 4 * C)
        P["Z"] = opb(NB * T, C); P["Y"] = f32(NB * T, C)
        # stem
        M0 = NB * T
        P["E"] = opb(M0, 3 * C)
        P["e"] = f32(M0, C)
        P["X"] = f32(M0, C); P["X1"] = f32(M0, C)
        P["Qin"] = opb(M0, C); P["Kin"] = opb(M0, C); P["Vin"] = opb(M0, C)
        P["Qp"] = f32(M0, C); P["Kp"] = f32(M0, C); P["Vp"] = f32(M0, C)
        if self.tc_attn:      # tcgen05 attention: q, k as operand rows, values transposed per item [NB*C, T]
            P["Qop"] = opb(M0, C); P["Kop"] = opb(M0, C); P["VT"] = opb(NB * C, T)
            P["q2op"] = opb(M0, C // 2); P["k2op"] = opb(M0, C // 2); P["VT2"] = opb(NB * (C // 2), T)
        P["AO"] = opb(M0, C)
        P["Hn"] = opb(M0, C); P["Hm"] = opb(M0, 4 * C)
        # pyramid / fusion
        P["P"] = [None] + [f32(NB * Tl[l], C) for l in range(1, L)]
        P["TDin"] = [opb(NB * Tl[l], 2 * C) for l in range(L - 1)]
        P["BUin"] = [None] + [opb(NB * Tl[l], 2 * C) for l in range(1, L)]
        P["u"] = [f32(NB * Tl[l], C) for l in range(L - 1)] + [None]
        P["o"] = [None] + [f32(NB * Tl[l], C) for l in range(1, L)]
        P["h"] = f32(M0, C)
        P["CAT"] = opb(M0, 3 * C)
        Ch = C // 2
        P["c"] = [f32(M0, Ch) for _ in range(3)]
        P["q2"] = opb(M0, Ch); P["k2"] = opb(M0, Ch); P["v2"] = opb(M0, Ch)
        P["qp2"] = f32(M0, Ch); P["kp2"] = f32(M0, Ch); P["vp2"] = f32(M0, Ch)
        P["ao2"] = opb(M0, Ch)
        P["c3i"] = opb(M0, 3 * Ch)
        P["gate"] = f32(M0, 8)
        P["gT"] = opb(NB * C, T)
        nG = (L - 1) * Ch
        P["G_td"] = f32(NB * C, nG); P["G_bu"] = f32(NB * C, nG)
        if self.tc_attn:
            P["G_td_op"] = opb(NB * C, nG); P["G_bu_op"] = opb(NB * C, nG)
        P["qm"] = f32(M0, C); P["g2"] = f32(M0, C)
        P["DS"] = opb(NB * Tl[1], 3 * C); P["dconv"] = f32(NB * Tl[1], C)
        # heads
        Mh = B * Ttot
        P["HIN"] = opb(Mh, 3 * 2 * C)
        P["hc1"] = f32(Mh, 2 * C)
        P["HC"] = opb(Mh, 3 * C); P["HR"] = opb(Mh, 3 * C)
        P["hc2"] = f32(Mh, C); P["hr2"] = f32(Mh, C)
        P["HC2"] = opb(Mh, 3 * C); P["HR2"] = opb(Mh, 3 * C)
        P["logits"] = f32(Mh, self.ncls); P["offsets"] = f32(Mh, 2 * self.ncls)
        P["m_heads"] = u8(Mh)
        rs = torch.empty(B, Ttot, dtype=torch.float32)
        for l in range(L):       # Scale_l per row (meta_archs.py:257), constant per plan
            rs[:, self.level_off[l]:self.level_off[l + 1]] = self.head_scales[l]
        P["rowscale"] = rs.reshape(Mh).to(dev)
        pts = []
        for l in range(L):
            s = float(self.model.fpn_strides[l])
            rr = self.model.reg_range[l]
            for t in range(Tl[l]):
                pts.append([t * s, float(rr[0]), float(rr[1]), s])
        P["points"] = torch.tensor(pts, dtype=torch.float32, device=dev)
        # decode / nms
        topk = self.model.test_pre_nms_topk
        P["cap"] = sum(min(topk, Tl[l] * self.ncls) for l in range(L))
        P["cand_segs"] = f32(B, P["cap"], 2); P["cand_scores"] = f32(B, P["cap"])
        P["cand_labels"] = ar.alloc((B, P["cap"]), torch.int32)
        Kd = self.model.test_max_seg_num
        P["out_segs"] = f32(B, Kd, 2); P["out_scores"] = f32(B, Kd)
        P["out_labels"] = ar.alloc((B, Kd), torch.int64)
        P["out_counts"] = ar.alloc((B,), torch.int32)
        P["nms_ws"] = ar.alloc((K.softnms_workspace_bytes(B, self.ncls, Kd),), torch.uint8)
        P["graph"] = None
        self._plans[(B, slot)] = P
        return P

    # ------------------------------------------------------------------------------ helpers
    def _gemm(self, groups, M, N, Kd, act=ACT_NONE, res_masked=False):
        K.gemm(groups, M, N, Kd, self.op, act, res_masked, self.backend, passes=self.stage_passes.get(self._stage, 0))

    def _tc_attn_for(self, T):
        """Tensor-core attention for this key length?  Short pyramid levels (T <= attn_simt_max_t) take the FP32 CUDA-core
        kernel: a (item, head) there is a <= 28 x 28 score matrix — one 128-query tcgen05 tile would be >= 78 % padding and
        hold a whole SM (168 registers x 320 threads, TMEM) for ~13 us, while the CUDA-core CTA is 128 threads / 53 KB."""
        return self.tc_attn and T > self.attn_simt_max_t

    def _qkv_outs(self, q32, k32, v32, qop, kop, vt, T, Cc, rows=None, items=None):
        """Output routing of the q / k / v projection GEMMs.  Tensor-core attention: q, k as operand rows and the
        values written TRANSPOSED per item ([item*Cc + channel, key]) straight from the GEMM epilogue."""
        sl = (lambda t: t) if rows is None else (lambda t: t[rows[0]:rows[1]])
        if self._tc_attn_for(T):
            vtv = vt if items is None else vt[items[0] * Cc:items[1] * Cc]
            return {"out_op": sl(qop)}, {"out_op": sl(kop)}, {"out_opT": vtv, "t_seg": T}
        return {"out_f32": sl(q32)}, {"out_f32": sl(k32)}, {"out_f32": sl(v32)}

    def _attend(self, q32, k32, v32, qop, kop, vt, kmask, out, nb, T, nh, hs, ws=None):
        """MaskedMHCA core (blocks.py:218-240): tcgen05 kernel on operand q/k + transposed values, or the CUDA-core one."""
        scale = 1.0 / math.sqrt(hs)
        if self._tc_attn_for(T):
            Cc = nh * hs
            # qmask = kmask: self-attention over one masked sequence — the projection that follows multiplies the rows of invalid
            # queries by the same mask (blocks.py:243), so tiles without a valid query are skipped (zero rows)
            K.attention_tc([{"q": qop, "k": kop, "vt": vt[:nb * Cc], "kmask": kmask, "qmask": kmask, "out": out}], nb, T, T, nh, hs, scale, self.op,
                           passes=self.stage_passes.get(self._stage, 0), workspace=ws)
        else:
            K.attention([{"q": q32, "k": k32, "v": v32, "kmask": kmask, "out": out}], nb, T, T, nh, hs, scale, self.op)

    def _mask(self, P, l, kind="true"):
        """u8 view of the level-l mask over the NB-stacked rows."""
        NB = 2 * P["B"]
        o = P["m_off"][l]
        buf = P["m_true"] if kind == "true" else P["m_up"]
        return buf[o:o + NB * self.Tl[l]]

    # ------------------------------------------------------------------------------ forward
    def _launch_all(self, P):
        self._launch_forward(P)
        self._launch_decode(P)
        self._launch_nms(P, P["vid_meta"])

    def _launch_forward(self, P):
        self._launch_trunk(P)
        if self.model.use_dependency:
            self._launch_dependency(P)
        self._launch_heads(P)

    def _launch_dependency(self, P):
        """Dependency_Block between the fusion outputs and the heads (multimodal_meta_archs.py:474-475), through the
        module-level kernels (``_fwd.dependency_block_forward``; un-fused, but captured into the CUDA graph with the rest).  The block's outputs replace
        the fusion outputs in place, in the same [visual rows | audio rows] split the heads' gather reads."""
        from . import _fwd
        B, C, L, Tl = P["B"], self.C, self.L, self.Tl
        o = P["o_final"]
        feats, masks = [], []
        for l in range(L):
            n = B * Tl[l]
            feats.append(_fwd.from_rows(torch.cat((o[l][:n], o[l][n:2 * n]), dim=1), B, Tl[l]))       # cat(V, A) (:469)
            masks.append(self._mask(P, l)[:n].view(B, 1, Tl[l]).bool())
        prev, _fwd.MODE = _fwd.MODE, self.mode
        try:
            outs, _ = self.model.dependency_block(feats, masks)
        finally:
            _fwd.MODE = prev
        for l in range(L):
            n = B * Tl[l]
            r = _fwd.to_rows(outs[l])
            o[l][:n].copy_(r[:, :C])
            o[l][n:2 * n].copy_(r[:, C:])

    def _launch_trunk(self, P):
        w, op, B, T, C, L = self.w, self.op, P["B"], self.T, self.C, self.L
        NB, Tl, Ttot = 2 * B, self.Tl, self.Ttot
        N1 = T + 1
        M0 = NB * T
        half = B * T                       # rows of one modality in NB-stacked level-0 matrices
        m0 = self._mask(P, 0)

        # ---- masks: true pyramid masks, up-sampled coarse masks, CLS-extended mask
        K.build_masks(P["mask_in"], P["m_true"], P["m_up"], P["m_cls"], P["m_heads"], NB, B, T, L)

        # ================================================================== Alignment (:1144-1207)
        self._stage = "alignment"
        K.transpose_cast(P["visual"], T, P["Xv"], B, 2048, T, op)        # [B,2048,T] -> [B*T,2048]
        K.transpose_cast(P["audio"], T, P["Xa"], B, 128, T, op)
        self._gemm([{"A": P["Xv"], "W": w["al.pv"], "bias": w["al.pv.b"], "out_f32": P["x0"][:half]}], half, C, 2048)
        self._gemm([{"A": P["Xa"], "W": w["al.pa"], "bias": w["al.pa.b"], "out_f32": P["x0"][half:]}], half, C, 128)
        K.align_embed(P["x0"], w["al.cls_token_video"], w["al.cls_token_text"], w["al.pos_v"], w["al.pos_a"],
                      w["al.type_video"], w["al.type_text"], P["F"], B, T, C)
        Ma = NB * N1
        hm = B * N1                        # rows of one modality in the token matrix
        F, F1 = P["F"], P["F1"]
        for _layer in range(self.model.alignment.num_layers):          # same weights twice (:1009)
            K.layernorm_rows([{"x": F, "w": w["al.n1.w"], "b": w["al.n1.b"], "out_op": P["Fn"]}], Ma, C, op)
            d = {"A": P["Fn"], "W": w["al.qkv"], "bias": w["al.qkv.b"], "out_f32": P["QKV"]}
            if self.tc_attn:
                d.update({"out_op": P["QKVop"], "out_opT": P["VTa"], "t_seg": N1, "t_col0": 2 * C, "t_ncols": C})
            self._gemm([d], Ma, 3 * C, C)
            qkv = P["QKV"]
            groups = []
            for g in range(2):
                own, oth = qkv[g * hm:(g + 1) * hm], qkv[(1 - g) * hm:(2 - g) * hm]
                if self.tc_attn:
                    oop = P["QKVop"][g * hm:(g + 1) * hm]
                    groups.append({"q": View(oop, 0, C), "k": View(oop, C, C), "vt": P["VTa"][g * B * C:(g + 1) * B * C],
                                   "kmask": P["m_cls"],      # no qmask here: nothing masks these rows before the embedding
                                   # convolution reads its first padded neighbour (blocks.py:36-61 masks the OUTPUT only), so the
                                   # padded query rows must hold what the reference computes, not zeros
                                   "q32": View(own, 0, C), "xk": View(oth, C, C), "xv": View(oth, 2 * C, C), "x_first": 1,
                                   "out": P["AOa"][g * hm:(g + 1) * hm]})
                else:
                    groups.append({"q": View(own, 0, C), "k": View(own, C, C), "v": View(own, 2 * C, C),
                                   "kmask": P["m_cls"], "xk": View(oth, C, C), "xv": View(oth, 2 * C, C), "x_first": 1,
                                   "out": P["AOa"][g * hm:(g + 1) * hm]})
            if self.tc_attn:
                K.attention_tc(groups, B, N1, N1, 8, C // 8, 1.0 / math.sqrt(C // 8), op, passes=self.stage_passes.get(self._stage, 0),
                               workspace=P.get("att_ws"))
            else:
                K.attention(groups, B, N1, N1, 8, C // 8, 1.0 / math.sqrt(C // 8), op)
            self._gemm([{"A": P["AOa"], "W": w["al.m"], "bias": w["al.m.b"], "res": F, "out_f32": F1}], Ma, C, C)
            K.layernorm_rows([{"x": F1[g * hm:(g + 1) * hm], "w": w[f"al.n2.{mod}.w"], "b": w[f"al.n2.{mod}.b"],
                               "out_op": P["Fn"][g * hm:(g + 1) * hm]} for g, mod in enumerate(("video", "text"))], hm, C, op)
            self._gemm([{"A": P["Fn"][g * hm:(g + 1) * hm], "W": w[f"al.fc1.{mod}"], "bias": w[f"al.fc1.{mod}.b"],
                         "out_op": P["Ha"][g * hm:(g + 1) * hm]} for g, mod in enumerate(("video", "text"))],
                       hm, 4 * C, C, act=ACT_GELU)
            self._gemm([{"A": P["Ha"][g * hm:(g + 1) * hm], "W": w[f"al.fc2.{mod}"], "bias": w[f"al.fc2.{mod}.b"],
                         "res": F1[g * hm:(g + 1) * hm], "out_f32": F[g * hm:(g + 1) * hm]}
                        for g, mod in enumerate(("video", "text"))], hm, C, 4 * C)
        # drop CLS, LN(residual + x), Linear -> ReLU -> LN (:1192-1198)
        K.layernorm_rows([{"x": F[g * hm:(g + 1) * hm], "x_seg_rows": T, "x_seg_stride": N1, "x_row_off": 1,
                           "add": P["x0"][g * half:(g + 1) * half], "w": w[f"al.nf.{mod}.w"], "b": w[f"al.nf.{mod}.b"],
                           "out_op": P["Z"][g * half:(g + 1) * half]} for g, mod in enumerate(("video", "text"))], half, C, op)
        self._gemm([{"A": P["Z"][g * half:(g + 1) * half], "W": w[f"al.fc.{mod}"], "bias": w[f"al.fc.{mod}.b"],
                     "out_f32": P["Y"][g * half:(g + 1) * half]} for g, mod in enumerate(("video", "text"))],
                   half, C, C, act=ACT_RELU)
        K.layernorm_rows([{"x": P["Y"][g * half:(g + 1) * half], "w": w[f"al.fcn.{mod}.w"], "b": w[f"al.fcn.{mod}.b"],
                           "edge": P["edge_T"][g * half:(g + 1) * half],
                           "out_im2col": P["E"][g * half:(g + 1) * half]} for g, mod in enumerate(("video", "text"))], half, C, op)

        # ================================================================== backbone stem (:771-807)
        self._stage = "backbone"
        X, X1 = P["X"], P["X1"]
        for i in range(2):
            self._gemm([{"A": P["E"][g * half:(g + 1) * half], "W": w[f"bb.embd{Xm}{i}"], "rowmask": m0[g * half:(g + 1) * half],
                         "out_f32": P["e"][g * half:(g + 1) * half]} for g, Xm in enumerate("VA")], half, C, 3 * C)
            grp = []
            for g, Xm in enumerate("VA"):
                d = {"x": P["e"][g * half:(g + 1) * half], "w": w[f"bb.embdn{Xm}{i}.w"], "b": w[f"bb.embdn{Xm}{i}.b"]}
                if i == 0:
                    d.update({"edge": P["edge_T"][g * half:(g + 1) * half], "out_im2col": P["E"][g * half:(g + 1) * half]})
                else:   # + pos_embd * mask (:794-802)
                    d.update({"post": w["bb.pe"], "post_rows": T, "rowmask": m0[g * half:(g + 1) * half],
                              "out_f32": X[g * half:(g + 1) * half]})
                grp.append(d)
            K.layernorm_rows(grp, half, C, op, act=ACT_GELU)
        n_stem = len(self.model.backbone.self_att_V)
        for i in range(n_stem):
            names = [f"bb.sa{Xm}{i}" for Xm in "VA"]
            # LN11/LN12 -> depthwise conv -> mask -> LN for q (from ln12), k, v (from ln11)   (blocks.py:314, :205-211)
            K.dwconv_ln([{"x": X[g * half:(g + 1) * half], "mask_out": m0[g * half:(g + 1) * half],
                          "pre": [(w[nm + ".ln11.w"], w[nm + ".ln11.b"]), (w[nm + ".ln12.w"], w[nm + ".ln12.b"])],
                          "outs": [{"dw": w[nm + ".attn.query.dw"], "ln_w": w[nm + ".attn.query.nw"], "ln_b": w[nm + ".attn.query.nb"],
                                    "src": 1, "out_op": P["Qin"][g * half:(g + 1) * half]},
                                   {"dw": w[nm + ".attn.key.dw"], "ln_w": w[nm + ".attn.key.nw"], "ln_b": w[nm + ".attn.key.nb"],
                                    "src": 0, "out_op": P["Kin"][g * half:(g + 1) * half]},
                                   {"dw": w[nm + ".attn.value.dw"], "ln_w": w[nm + ".attn.value.nw"], "ln_b": w[nm + ".attn.value.nb"],
                                    "src": 0, "out_op": P["Vin"][g * half:(g + 1) * half]}]}
                         for g, nm in enumerate(names)], B, T, 1, C, op)
            grp = []
            for g, nm in enumerate(names):
                outs = self._qkv_outs(P["Qp"], P["Kp"], P["Vp"], P.get("Qop"), P.get("Kop"), P.get("VT"), T, C,
                                      rows=(g * half, (g + 1) * half), items=(g * B, (g + 1) * B))
                for xin, o, key in ((P["Qin"], outs[0], "query"), (P["Kin"], outs[1], "key"), (P["Vin"], outs[2], "value")):
                    grp.append(dict({"A": xin[g * half:(g + 1) * half], "W": w[f"{nm}.attn.{key}"], "bias": w[f"{nm}.attn.{key}.b"]}, **o))
            self._gemm(grp, half, C, C)
            hs = C // self.n_head
            self._attend(P["Qp"], P["Kp"], P["Vp"], P.get("Qop"), P.get("Kop"), P.get("VT"), m0, P["AO"], NB, T, self.n_head, hs,
                         ws=P.get("att_ws"))
            # out = x*mask + scale_attn * (proj(att)*mask)   (blocks.py:243, :316)
            self._gemm([{"A": P["AO"][g * half:(g + 1) * half], "W": w[nm + ".attn.proj"], "bias": w[nm + ".attn.proj.b"],
                         "rowmask": m0[g * half:(g + 1) * half], "res": X[g * half:(g + 1) * half], "colscale": w[nm + ".sa"],
                         "out_f32": X1[g * half:(g + 1) * half]} for g, nm in enumerate(names)], half, C, C, res_masked=True)
            K.layernorm_rows([{"x": X1[g * half:(g + 1) * half], "w": w[nm + ".ln2.w"], "b": w[nm + ".ln2.b"],
                               "out_op": P["Hn"][g * half:(g + 1) * half]} for g, nm in enumerate(names)], half, C, op)
            self._gemm([{"A": P["Hn"][g * half:(g + 1) * half], "W": w[nm + ".mlp0"], "bias": w[nm + ".mlp0.b"],
                         "out_op": P["Hm"][g * half:(g + 1) * half]} for g, nm in enumerate(names)], half, 4 * C, C, act=ACT_GELU)
            # out = out + scale_mlp * (mlp(ln2(out)) * mask)   (blocks.py:318)
            grp = []
            for g, nm in enumerate(names):
                d = {"A": P["Hm"][g * half:(g + 1) * half], "W": w[nm + ".mlp3"], "bias": w[nm + ".mlp3.b"],
                     "rowmask": m0[g * half:(g + 1) * half], "res": X1[g * half:(g + 1) * half], "colscale": w[nm + ".sm"],
                     "out_f32": X[g * half:(g + 1) * half]}
                if i == n_stem - 1:        # level-0 pyramid feature straight into the top-down concat operand
                    d["out_op"] = View(P["TDin"][0], C, C).rows(g * half, (g + 1) * half)
                grp.append(d)
            self._gemm(grp, half, C, 4 * C)

        # ================================================================== pyramid (:813-829, shared weights)
        feats = [X] + P["P"][1:]
        for l in range(L - 1):
            o = {"dw": w[f"bb.down{l}.dw"], "ln_w": w[f"bb.down{l}.w"], "ln_b": w[f"bb.down{l}.b"], "out_f32": feats[l + 1]}
            o["out_op"] = View(P["TDin"][l + 1], C, C) if l + 1 < L - 1 else View(P["BUin"][L - 1], C, C)
            K.dwconv_ln([{"x": feats[l], "mask_out": self._mask(P, l + 1), "outs": [o]}], NB, Tl[l], 2, C, op)

        # ================================================================== fusion, both passes as one 2B batch (:552-619)
        self._stage = "fusion"
        # guide of item i = stem output of the other modality: rows rolled by B*T
        K.transpose_cast(X[half:], C, P["gT"][:B * C], B, T, C, op)      # [B,T,C] -> [B*C, T]
        K.transpose_cast(X[:half], C, P["gT"][B * C:], B, T, C, op)
        nG = (L - 1) * (C // 2)
        self._gemm([dict({"A": P["gT"], "W": w["fu.td.gfc"], "bias": w["fu.td.gfc.b"]},
                         **({"out_op": P["G_td_op"]} if self.tc_attn else {"out_f32": P["G_td"]}))], NB * C, nG, T)
        u = P["u"][:L - 1] + [feats[L - 1]]                                  # u_5 = p_5
        for idx in range(L - 1, 0, -1):
            l = idx - 1                                                      # output level
            K.rowcopy([{"src": u[idx], "dst": View(P["TDin"][l], 0, C), "nseg": NB, "seg_len_in": Tl[idx],
                        "seg_len_out": Tl[l], "num": 1, "den": 2, "C": C}], op)   # nearest x2 up-sample (:565-566)
            out_op = View(P["BUin"][l], C, C) if l >= 1 else None
            self._csp(P, f"fu.td{L - 1 - idx}", P["TDin"][l], P["G_td"], (L - 1 - idx) * (C // 2), self.td_heads[L - 1 - idx],
                      self._mask(P, l, "up"), l, u[l], out_op)
        # guide enhancement (:591-600): pooled top-down outputs -> match projection -> query of text_enhancer
        K.pool_match(u[0], u[1], u[2], Tl[0], Tl[1], Tl[2], w["fu.match.w"], w["fu.match.b"], P["qm"], NB, C, T, 4)
        te = "fu.te"
        kv_out = lambda key, buf: {"dw": w[f"{te}.{key}.dw"], "ln_w": w[f"{te}.{key}.nw"], "ln_b": w[f"{te}.{key}.nb"], "out_op": buf}
        K.dwconv_ln([{"x": X[half:], "mask_out": m0[:half], "outs": [kv_out("key", P["Kin"][:half]), kv_out("value", P["Vin"][:half])]},
                     {"x": X[:half], "mask_out": m0[half:], "outs": [kv_out("key", P["Kin"][half:]), kv_out("value", P["Vin"][half:])]}],
                    B, T, 1, C, op)
        K.dwconv_ln([{"x": P["qm"], "mask_out": m0, "outs": [kv_out("query", P["Qin"])]}], NB, T, 1, C, op)
        outs = self._qkv_outs(P["Qp"], P["Kp"], P["Vp"], P.get("Qop"), P.get("Kop"), P.get("VT"), T, C)
        self._gemm([dict({"A": xin, "W": w[f"{te}.{key}"], "bias": w[f"{te}.{key}.b"]}, **o)
                    for xin, o, key in ((P["Qin"], outs[0], "query"), (P["Kin"], outs[1], "key"), (P["Vin"], outs[2], "value"))],
                   M0, C, C)
        hs = C // 4
        self._attend(P["Qp"], P["Kp"], P["Vp"], P.get("Qop"), P.get("Kop"), P.get("VT"), m0, P["AO"], NB, T, 4, hs, ws=P.get("att_ws"))
        self._gemm([{"A": P["AO"], "W": w[f"{te}.proj"], "bias": w[f"{te}.proj.b"], "rowmask": m0, "out_f32": P["g2"]}], M0, C, C)
        K.transpose_cast(P["g2"], C, P["gT"], NB, T, C, op)
        self._gemm([dict({"A": P["gT"], "W": w["fu.bu.gfc"], "bias": w["fu.bu.gfc.b"]},
                         **({"out_op": P["G_bu_op"]} if self.tc_attn else {"out_f32": P["G_bu"]}))], NB * C, nG, T)
        # bottom-up (:602-612)
        o = [u[0]] + P["o"][1:]
        for l in range(L - 1):
            Mn = NB * Tl[l + 1]
            mt = self._mask(P, l + 1)
            K.rowcopy([{"src": o[l], "dst": P["DS"], "nseg": NB, "seg_len_in": Tl[l], "seg_len_out": Tl[l + 1],
                        "num": 2, "den": 1, "ntaps": 3, "tap_stride": C, "C": C}], op)        # stride-2 im2col
            self._gemm([{"A": P["DS"], "W": w["fu.ds"], "bias": w["fu.ds.b"], "rowmask": mt, "out_f32": P["dconv"]}], Mn, C, 3 * C)
            K.layernorm_rows([{"x": P["dconv"], "w": w["fu.dsn.w"], "b": w["fu.dsn.b"], "out_op": View(P["BUin"][l + 1], 0, C)}],
                             Mn, C, op, act=ACT_SILU)
            self._csp(P, f"fu.bu{l}", P["BUin"][l + 1], P["G_bu"], l * (C // 2), self.bu_heads[l], mt, l + 1, o[l + 1], None)

        P["o_final"] = o

    def _launch_heads(self, P):
        # ================================================================== heads (meta_archs.py:166-178, :245-259)
        w, op, B, C, L = self.w, self.op, P["B"], self.C, self.L
        Tl, Ttot = self.Tl, self.Ttot
        o = P["o_final"]
        self._stage = "heads"
        Mh = B * Ttot
        jobs = []
        for l in range(L):
            for m in range(2):         # feats_AV = cat(V, A) along channels (:469)
                jobs.append({"src": o[l][m * B * Tl[l]:(m + 1) * B * Tl[l]], "dst": View(P["HIN"], m * C, C), "nseg": B,
                             "seg_len_in": Tl[l], "seg_len_out": Tl[l], "dst_seg_stride": Ttot, "dst_row_off": self.level_off[l],
                             "ntaps": 3, "tap_stride": 2 * C, "C": C})
        K.rowcopy(jobs, op)
        mh = P["m_heads"]
        self._gemm([{"A": P["HIN"], "W": w["hd.c0"], "rowmask": mh, "out_f32": P["hc1"]}], Mh, 2 * C, 6 * C)
        K.layernorm_rows([{"x": View(P["hc1"], 0, C), "w": w["hd.cls.n0.w"], "b": w["hd.cls.n0.b"], "edge": P["edge_heads"], "out_im2col": P["HC"]},
                          {"x": View(P["hc1"], C, C), "w": w["hd.reg.n0.w"], "b": w["hd.reg.n0.b"], "edge": P["edge_heads"], "out_im2col": P["HR"]}],
                         Mh, C, op, act=ACT_RELU)
        self._gemm([{"A": P["HC"], "W": w["hd.cls.c1"], "rowmask": mh, "out_f32": P["hc2"]},
                    {"A": P["HR"], "W": w["hd.reg.c1"], "rowmask": mh, "out_f32": P["hr2"]}], Mh, C, 3 * C)
        K.layernorm_rows([{"x": P["hc2"], "w": w["hd.cls.n1.w"], "b": w["hd.cls.n1.b"], "edge": P["edge_heads"], "out_im2col": P["HC2"]},
                          {"x": P["hr2"], "w": w["hd.reg.n1.w"], "b": w["hd.reg.n1.b"], "edge": P["edge_heads"], "out_im2col": P["HR2"]}],
                         Mh, C, op, act=ACT_RELU)
        self._gemm([{"A": P["HC2"], "W": w["hd.cls.out"], "bias": w["hd.cls.out.b"], "rowmask": mh, "out_f32": P["logits"]}],
                   Mh, self.ncls, 3 * C)
        # offsets = relu(scale_l * (conv * mask))   (:256-257)
        self._gemm([{"A": P["HR2"], "W": w["hd.reg.out"], "bias": w["hd.reg.out.b"], "rowmask": mh, "rowscale": P["rowscale"],
                     "out_f32": P["offsets"]}], Mh, 2 * self.ncls, 3 * C, act=ACT_RELU)

    def _launch_decode(self, P):
        """Candidate decode (meta_archs.py:745-817) of the head outputs in P."""
        md = self.model
        B, mh = P["B"], P["m_heads"]
        K.decode(P["logits"], P["offsets"], mh, P["points"], self.level_off, B, self.ncls, md.class_aware,
                 md.test_pre_nms_thresh, md.test_pre_nms_topk, md.test_duration_thresh, P["cand_segs"], P["cand_scores"],
                 P["cand_labels"], P["cap"])

    def _launch_nms(self, P, vid_meta):
        """Per-class soft-NMS + merge + seconds (meta_archs.py:819-875, libs/utils/nms.py:103-190) of P's candidates."""
        md = self.model
        B, Ttot = P["B"], self.Ttot
        from .utils.nms import nms_method_code
        method = nms_method_code(md.test_nms_method, md.test_multiclass_nms)
        K.softnms_batched(P["cand_segs"], P["cand_scores"], P["cand_labels"], B, P["cap"], self.ncls, md.test_iou_threshold,
                          md.test_nms_sigma, md.test_min_score, method, md.test_max_seg_num, Ttot, vid_meta,
                          P["out_segs"], P["out_scores"], P["out_labels"], P["out_counts"], P["nms_ws"])

    def _csp(self, P, name, xin, G, g_off, heads, mask, l, out_f32, out_op):
        """MaxSigmoidCSPLayerWithTwoConv.forward (multimodal_backbones.py:243-256) at pyramid level l."""
        w, op, C = self.w, self.op, self.C
        NB = 2 * P["B"]
        Tl = self.Tl[l]
        M = NB * Tl
        Ch = C // 2
        CAT = P["CAT"][:M]
        h = P["h"][:M]
        # main 1x1 conv (+bias) * mask; split halves h_a | h_b live in `h`, operand copy in CAT[:, 0:C]
        self._gemm([{"A": xin, "W": w[name + ".main"], "bias": w[name + ".main.b"], "rowmask": mask, "out_f32": h,
                     "out_op": View(CAT, 0, C)}], M, C, 2 * C)
        x = View(h, Ch, Ch)
        q2, k2, v2 = P["q2"][:M], P["k2"][:M], P["v2"][:M]
        qp, kp, vp = P["qp2"][:M], P["kp2"][:M], P["vp2"][:M]
        q2op = P["q2op"][:M] if self.tc_attn else None
        k2op = P["k2op"][:M] if self.tc_attn else None
        ao = P["ao2"][:M]
        for j in range(3):
            nm = f"{name}.blk{j}"
            K.dwconv_ln([{"x": x, "mask_out": mask,
                          "outs": [{"dw": w[f"{nm}.{key}.dw"], "ln_w": w[f"{nm}.{key}.nw"], "ln_b": w[f"{nm}.{key}.nb"], "out_op": buf}
                                   for key, buf in (("query", q2), ("key", k2), ("value", v2))]}], NB, Tl, 1, Ch, op)
            outs = self._qkv_outs(qp, kp, vp, q2op, k2op, P.get("VT2"), Tl, Ch)
            self._gemm([dict({"A": a, "W": w[f"{nm}.{key}"], "bias": w[f"{nm}.{key}.b"]}, **o)
                        for a, o, key in ((q2, outs[0], "query"), (k2, outs[1], "key"), (v2, outs[2], "value"))], M, Ch, Ch)
            self._attend(qp, kp, vp, q2op, k2op, P.get("VT2"), mask, ao, NB, Tl, 4, Ch // 4, ws=P.get("att_ws"))
            cj = P["c"][j][:M]
            self._gemm([{"A": ao, "W": w[f"{nm}.proj"], "bias": w[f"{nm}.proj.b"], "rowmask": mask, "out_f32": cj,
                         "out_op": View(CAT, C + j * Ch, Ch)}], M, Ch, Ch)
            x = cj
        c3 = P["c"][2][:M]
        hc = Ch // heads
        gate = P["gate"].view(-1)[:M * heads].view(M, heads)
        if self.tc_attn and C == 512:      # tcgen05 gate: c_3's operand copy is the CAT window written by block 2's projection
            K.maxsig_gate_tc(CAT, C + 2 * Ch, P["G_td_op"] if G is P["G_td"] else P["G_bu_op"], g_off, w[name + ".hb"], gate,
                             NB, Tl, C, heads, hc, op, passes=self.stage_passes.get(self._stage, 0))
        else:
            K.maxsig_gate(c3, View(G, g_off, Ch), w[name + ".hb"], gate, NB, Tl, C, heads, hc)
        if self.implicit_c3 and self.tc_attn:
            # implicit k = 3 convolution: the GEMM reads c_3's PLAIN operand (the CAT window block 2's projection wrote) through a
            # 4-D tensor map, tap t of a tile from rows t0 + t - 1 with TMA zero fill outside the item — no im2col operand, no copy
            # launch; same accumulation order as the im2col GEMM (test_implicit_conv3_equals_im2col_gemm)
            self._gemm([{"A": View(CAT, C + 2 * Ch, Ch), "conv_T": Tl, "W": w[name + ".proj"], "bias": w[name + ".proj.b"], "rowmask": mask,
                         "gate": gate, "gate_groups": heads, "gate_width": hc, "out_op": View(CAT, C + 3 * Ch, Ch)}], M, Ch, 3 * Ch)
        else:
            K.rowcopy([{"src": c3, "dst": P["c3i"][:M], "nseg": NB, "seg_len_in": Tl, "seg_len_out": Tl, "ntaps": 3,
                        "tap_stride": Ch, "C": Ch}], op)
            self._gemm([{"A": P["c3i"][:M], "W": w[name + ".proj"], "bias": w[name + ".proj.b"], "rowmask": mask, "gate": gate,
                         "gate_groups": heads, "gate_width": hc, "out_op": View(CAT, C + 3 * Ch, Ch)}], M, Ch, 3 * Ch)
        d = {"A": CAT, "W": w[name + ".final"], "bias": w[name + ".final.b"], "rowmask": mask, "out_f32": out_f32}
        if out_op is not None:
            d["out_op"] = out_op
        self._gemm([d], M, C, 3 * C)

    # ------------------------------------------------------------------------------ public
    @torch.no_grad()
    def run(self, visual: torch.Tensor, audio: torch.Tensor, mask: torch.Tensor, vid_meta: torch.Tensor,
            overlap_nms: bool = False, slot: int = 0):
        """visual [B,2048,T] f32, audio [B,128,T] f32, mask [B,1,T] bool, vid_meta [B,4] f32
        (feat_stride, feat_num_frames, fps, duration) — any device; copied into the plan's static inputs.
        Returns the plan dict (device-resident outputs: out_segs/out_scores/out_labels/out_counts, logits, offsets).

        overlap_nms=False: everything is enqueued on the current stream (one CUDA graph); ``slot`` must be 0.
        overlap_nms=True : streaming mode for back-to-back batches.  Plan ``slot`` owns a forward stream and an NMS stream:
        the input copies, the CUDA graph of the forward and the decode run on the former, the soft-NMS kernel (one CTA
        per video: 16 of 148 SMs, a <=100-round dependency chain) on the latter, so it overlaps the next batch's forward.
        Alternating two slots keeps two batches in flight: the latency-bound short-pyramid-level kernels of one (a few
        CTAs each, ~40 % of a step) run under the big GEMMs of the other.  The caller's stream only waits for the input
        copies (it may recycle its input tensors afterwards); consumers of out_* order themselves after
        ``plan["ev_nms"]`` or enqueue on ``plan["nms_stream"]``."""
        B = visual.shape[0]
        if visual.shape[2] != self.T or audio.shape[2] != self.T:
            # the reference pads every batch to max_seq_len (data_utils.py:170-176) and, for longer videos, to a multiple of
            # max_div_factor — the static plan is built for max_seq_len only
            raise NotImplementedError(f"HotPathEngine: sequence length {visual.shape[2]} != max_seq_len {self.T}; pad to max_seq_len "
                                      "(videos longer than max_seq_len are not supported by the fused plan)")
        assert overlap_nms or slot == 0
        P = self._plan(B, slot)
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream()
            if not overlap_nms:
                if P.get("ev_nms") is not None:
                    cur.wait_event(P["ev_nms"])                   # a streamed step of this plan may still be running
                    P["ev_nms"] = None
                P["visual"].copy_(visual, non_blocking=True)
                P["audio"].copy_(audio, non_blocking=True)
                P["mask_in"].copy_(mask.reshape(B, self.T), non_blocking=True)
                P["vid_meta"].copy_(vid_meta, non_blocking=True)
                if not self.use_graph:
                    self._launch_all(P)
                else:
                    if P["graph"] is None:
                        self._launch_all(P)                     # warm-up (sets function attributes, fills constants)
                        cur.synchronize()
                        n0 = K.launch_count()
                        g = torch.cuda.CUDAGraph()
                        # latency plans (batch <= 2, synchronous path): programmatic dependent launch captured into the graph
                        # — the next kernel's prologue overlaps the previous kernel's tail (2.83 -> 2.64 ms at batch 1); not
                        # for the streamed plans, where early-launched CTAs take SM resources from the other batches
                        lib = _cabi.load(self.op)
                        pdl = B <= 2 and os.environ.get("UNAV_PDL") is None
                        if pdl:
                            lib.unav_set_pdl(1)
                        try:
                            with torch.cuda.graph(g):
                                self._launch_all(P)
                        finally:
                            if pdl:
                                lib.unav_set_pdl(-1)
                        P["launches_per_step"] = K.launch_count() - n0
                        P["graph"] = g
                    P["graph"].replay()
                return P
            # ---- streaming mode
            if "meta2" not in P:
                P["meta2"] = P["arena"].alloc((2, B, 4), torch.float32)
                P["ev_dec"] = torch.cuda.Event()
                P["step"] = 0
                P["fwd_stream"] = torch.cuda.Stream(self.dev)
                P["nms_stream"] = torch.cuda.Stream(self.dev)
            if self.use_graph and P.get("graph_fwd") is None:
                self._launch_all(P)                             # warm-up on the caller's stream
                torch.cuda.synchronize(self.dev)
                n0 = K.launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_forward(P)
                P["launches_per_step"] = K.launch_count() - n0 + 2
                P["graph_fwd"] = g
            fs, ns = P["fwd_stream"], P["nms_stream"]
            meta = P["meta2"][P["step"] & 1]
            P["step"] += 1
            ev_in = torch.cuda.Event()
            ev_in.record(cur)
            fs.wait_event(ev_in)                                # the inputs are ready in the caller's stream order
            with torch.cuda.stream(fs):
                P["visual"].copy_(visual, non_blocking=True)
                P["audio"].copy_(audio, non_blocking=True)
                P["mask_in"].copy_(mask.reshape(B, self.T), non_blocking=True)
                meta.copy_(vid_meta, non_blocking=True)
                ev_cp = torch.cuda.Event()
                ev_cp.record(fs)
                if self.use_graph:
                    P["graph_fwd"].replay()
                else:
                    self._launch_forward(P)
                if P.get("ev_nms") is not None:
                    fs.wait_event(P["ev_nms"])                  # this slot's previous NMS has consumed the candidates
                self._launch_decode(P)
                P["ev_dec"].record(fs)
            cur.wait_event(ev_cp)                               # caller-side tensors may be recycled after the copies
            with torch.cuda.stream(ns):
                ns.wait_event(P["ev_dec"])
                self._launch_nms(P, meta)
                ev = torch.cuda.Event()
                ev.record(ns)
                P["ev_nms"] = ev
        return P

    def capture_traced(self, B: int):
        """A second CUDA graph of the same launches with an event-record node on each side of every kernel
        (bench.py's roofline leg).  Returns (graph, records); after ``graph.replay()`` + synchronize,
        ``kernels.read_trace(records)`` gives the per-launch durations of that replay."""
        P = self._plan(B)
        with torch.cuda.device(self.dev):
            self._launch_all(P)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            K.start_trace(external=True)
            try:
                with torch.cuda.graph(g):
                    self._launch_all(P)
            finally:
                rec = K.stop_trace(read=False)
        return g, rec
