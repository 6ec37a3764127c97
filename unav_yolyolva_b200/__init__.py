"""B200-native implementation of the UnAV_yolyolVA inference hot path (PtTransformer forward + decode +
per-class temporal soft-NMS) behind the reference's module / operator surface.  See DESIGN.md."""
__version__ = "0.1.0"
