from torch import nn


class BaseModule(nn.Module):
    """Accepts and ignores ``init_cfg`` like mmengine's BaseModule."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
