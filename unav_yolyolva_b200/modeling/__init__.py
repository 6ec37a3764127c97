"""``libs.modeling`` surface of the reference (/root/reference/libs/modeling/__init__.py:1-12)."""
from .blocks import MaskedConv1D, MaskedMHCA, LayerNorm, TransformerBlock, Scale, AffineDropPath
from .models import make_multimodal_backbone, make_multimodal_meta_arch, make_dependency_block
from . import multimodal_backbones  # noqa: F401  (registers convTransformer)
from . import dependency_block      # noqa: F401  (registers DependencyBlock)
from . import multimodal_meta_archs  # noqa: F401  (registers LocPointTransformer)

__all__ = ["MaskedConv1D", "MaskedMHCA", "LayerNorm", "TransformerBlock", "Scale", "AffineDropPath",
           "make_multimodal_backbone", "make_multimodal_meta_arch", "make_dependency_block"]
