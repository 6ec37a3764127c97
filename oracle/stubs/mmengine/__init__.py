"""Import stub (test infrastructure only): the reference imports
``mmengine.model.BaseModule`` purely as a base class
(/root/reference/libs/modeling/multimodal_backbones.py:12,51,118). mmengine is
not installed in this image, so the oracle harness puts this directory on
``sys.path`` before importing the reference."""
