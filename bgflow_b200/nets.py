"""Conditioner networks (host-side parameter owners).

Mirrors bgflow/nn/dense.py:9-54 (DenseNet, MeanFreeDenseNet) and bgflow/nn/periodic.py:7-37
(WrapPeriodic).  Parameters are ordinary ``torch.nn.Linear`` modules inside ``_layers`` (an
``nn.Sequential`` interleaving Linear and activation modules), so ``state_dict`` keys and
shapes are the same as the reference's.  Inside a coupling block these modules are never
*executed* layer by layer: the fused kernel reads their weights (see
``transformers._net_spec``).  Calling them directly evaluates the plain PyTorch definition on
whatever device the input lives on (used for generic conditioners and for the autograd path).
"""

import numpy as np
import torch

__all__ = ["DenseNet", "MeanFreeDenseNet", "WrapPeriodic"]


def _is_seq(x):
    return isinstance(x, (list, tuple))


class DenseNet(torch.nn.Module):
    """Multi-layer perceptron ``n_units[0] -> ... -> n_units[-1]`` (dense.py:10-45).

    ``activation`` is a module applied after every hidden layer (or one module per hidden
    layer); no activation follows the output layer.
    """

    def __init__(self, n_units, activation=None, weight_scale=1.0, bias_scale=0.0):
        super().__init__()
        n_layers = len(n_units) - 1
        if _is_seq(activation):
            assert len(activation) == len(n_units) - 2
        mods = []
        for i in range(n_layers):
            lin = torch.nn.Linear(n_units[i], n_units[i + 1])
            lin.weight.data *= weight_scale
            if bias_scale > 0.0:
                lin.bias.data = torch.Tensor(lin.bias.data).uniform_() * bias_scale
            mods.append(lin)
            if i < n_layers - 1 and activation is not None:
                mods.append(activation[i] if _is_seq(activation) else activation)
        self._layers = torch.nn.Sequential(*mods)

    def forward(self, x):
        return self._layers(x)


class MeanFreeDenseNet(DenseNet):
    """dense.py:51-54"""

    def forward(self, x):
        y = self._layers(x)
        return y - y.mean(dim=1, keepdim=True)


class WrapPeriodic(torch.nn.Module):
    """Feed periodic inputs to ``net`` as (cos, sin) pairs (periodic.py:7-37): the net sees
    ``[cos(2 pi (x_c - left)/(right - left)) .., sin(..) .., x_others ..]``."""

    def __init__(self, net, left=0.0, right=1.0, indices=slice(None)):
        super().__init__()
        self.net = net
        self.left = left
        self.right = right
        self.indices = indices

    def periodic_indices(self, width):
        return np.arange(width)[self.indices]

    def forward(self, x):
        idx = self.periodic_indices(x.shape[-1])
        others = np.setdiff1d(np.arange(x.shape[-1]), idx)
        arg = 2 * np.pi * (x[..., idx] - self.left) / (self.right - self.left)
        return self.net.forward(torch.cat([torch.cos(arg), torch.sin(arg), x[..., others]], dim=-1))
