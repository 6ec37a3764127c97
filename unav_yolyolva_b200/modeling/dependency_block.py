"""``DependencyBlock`` registry entry (/root/reference/libs/modeling/dependency_block.py:6-70).

The block is disabled in both reference configs (``use_dependency: False``,
configs/avel_unav100.yaml:15); SURVEY.md §8(f) rank 3.  Same constructor and parameter names, so a
``use_dependency=True`` checkpoint loads; the forward runs on the CUDA kernels through the module-level path
(``_fwd.dependency_block_forward``: GEMM / dwconv+LN / tensor-core attention kernels with torch views in between — not
the fused engine, and not tuned: it doubles the model's FLOPs).
"""
from torch import nn

from .blocks import MaskedConv1D, TransformerBlock
from .models import register_dependency_block


@register_dependency_block("DependencyBlock")
class Dependency_Block(nn.Module):
    def __init__(self, in_channel, n_embd, n_embd_ks, num_classes, path_pdrop, n_head=1):
        super().__init__()
        self.num_classes = num_classes
        self.relu = nn.ReLU(inplace=True)
        self.feature_expand = MaskedConv1D(in_channel, n_embd * num_classes, n_embd_ks, stride=1,
                                           padding=n_embd_ks // 2, bias=False)
        self.cooccur_branch = TransformerBlock(n_embd, n_head, n_hidden=n_embd, path_pdrop=path_pdrop)
        self.temporal_branch = TransformerBlock(n_embd, n_head, n_hidden=n_embd, path_pdrop=path_pdrop)
        self.feature_squeeze = MaskedConv1D(n_embd * num_classes, in_channel, n_embd_ks, stride=1,
                                            padding=n_embd_ks // 2, bias=False)

    def forward(self, fpn_feats, fpn_masks):
        from .. import _fwd
        return _fwd.dependency_block_forward(self, fpn_feats, fpn_masks)
