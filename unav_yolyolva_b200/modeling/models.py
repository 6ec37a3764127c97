"""Name -> class registries for the hot path's module surface.

Mirrors the plugin API of the reference (/root/reference/libs/modeling/models.py:3-35):
three dictionaries filled by decorators and three ``make_*`` builders.  The registered
names are the reference's: ``convTransformer`` (multimodal_backbones.py:625),
``DependencyBlock`` (dependency_block.py:6) and ``LocPointTransformer``
(multimodal_meta_archs.py:262).
"""

multimodal_backbones = {}
dependency_blocks = {}
multimodal_meta_archs = {}


def _registrar(table):
    def register(name):
        def decorator(cls):
            table[name] = cls
            return cls
        return decorator
    return register


register_multimodal_backbone = _registrar(multimodal_backbones)
register_dependency_block = _registrar(dependency_blocks)
register_multimodal_meta_arch = _registrar(multimodal_meta_archs)


def make_multimodal_backbone(name, **kwargs):
    return multimodal_backbones[name](**kwargs)


def make_dependency_block(name, **kwargs):
    return dependency_blocks[name](**kwargs)


def make_multimodal_meta_arch(name, **kwargs):
    return multimodal_meta_archs[name](**kwargs)
