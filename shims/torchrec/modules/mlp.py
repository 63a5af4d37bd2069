"""torchrec.modules.mlp.MLP as the dense arches use it (baselines/models/dlrm.py:130,236): a stack of Linear layers,
each followed by the activation."""
from typing import Callable, List, Union

import torch
import torch.nn as nn

_ACTIVATIONS = {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}


class Perceptron(nn.Module):
    def __init__(self, in_size: int, out_size: int, bias: bool = True,
                 activation: Union[str, Callable[[torch.Tensor], torch.Tensor]] = torch.relu, device=None):
        super().__init__()
        self._linear = nn.Linear(in_size, out_size, bias=bias, device=device)
        self._activation_fn = _ACTIVATIONS[activation] if isinstance(activation, str) else activation

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._activation_fn(self._linear(x))


class MLP(nn.Module):
    def __init__(self, in_size: int, layer_sizes: List[int], bias: bool = True,
                 activation: Union[str, Callable[[torch.Tensor], torch.Tensor]] = torch.relu, device=None):
        super().__init__()
        sizes = [in_size] + list(layer_sizes)
        self._mlp = nn.Sequential(*[Perceptron(sizes[i], sizes[i + 1], bias=bias, activation=activation, device=device)
                                    for i in range(len(layer_sizes))])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._mlp(x)
