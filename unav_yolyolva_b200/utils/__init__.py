"""``libs.utils`` operator surface on the hot path (/root/reference/libs/utils/__init__.py:1)."""
from .metrics import ANETdetection
from .nms import batched_nms

__all__ = ["batched_nms", "ANETdetection"]
