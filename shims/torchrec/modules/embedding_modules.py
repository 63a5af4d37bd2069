"""Named only in type annotations of baselines/models/dlrm.py (the torchrec comparison harness is out of scope)."""
import torch.nn as nn


class EmbeddingBagCollection(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torchrec EmbeddingBagCollection is not available in the shim")
