"""Residual-MLP worm position predictor: module classes with the reference's attribute layout
(wtracker/neural/mlp.py: WormPredictor :31-48, MLPLayer :51-89, MlpBlock :92-141, RMLP :144-188) so
that the committed whole-module checkpoints (``torch.save(model)``, neural/training.py:142)
unpickle into them, plus the loader that maps the pickled ``wtracker.*`` class paths here.

These torch modules are containers for weights; inference on the hot path goes through
``wtracker_b200.neural.engine.ResMLPEngine`` (CUDA), not through ``forward``.
"""

from __future__ import annotations

import pickle
from typing import Sequence, Union

import torch
from torch import Tensor, nn

from wtracker_b200.neural.config import IOConfig

ACTIVATIONS = {
    "relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "softmax": nn.Softmax, "logsoftmax": nn.LogSoftmax,
    "lrelu": nn.LeakyReLU, "none": nn.Identity, None: nn.Identity,
}
ACTIVATION_DEFAULT_KWARGS = {"softmax": dict(dim=1), "logsoftmax": dict(dim=1)}


def _activation(act: Union[str, nn.Module, None]) -> nn.Module:
    if isinstance(act, nn.Module):
        return act
    return ACTIVATIONS[act](**ACTIVATION_DEFAULT_KWARGS.get(act, {}))


class WormPredictor(nn.Module):
    """A model plus the IOConfig naming the frames it consumes / predicts."""

    def __init__(self, model: nn.Module, io_config: IOConfig):
        super().__init__()
        self.io_config: IOConfig = io_config
        self.model: nn.Module = model

    def forward(self, x: Tensor) -> Tensor:
        return self.model(x)


class MLPLayer(nn.Module):
    """Linear -> (BatchNorm1d) -> activation, stored as ``mlp_layer`` (a Sequential)."""

    def __init__(self, in_dim: int, out_dim: int, nonlin: Union[str, nn.Module], batch_norm: bool = True) -> None:
        super().__init__()
        parts: list[nn.Module] = [nn.Linear(in_dim, out_dim)]
        if batch_norm and nonlin not in ["none", None]:
            parts.append(nn.BatchNorm1d(out_dim))
        parts.append(_activation(nonlin))
        self.mlp_layer = nn.Sequential(*parts)

    def forward(self, x: Tensor) -> Tensor:
        return self.mlp_layer(x.reshape(x.size(0), -1))


class MlpBlock(nn.Module):
    """A chain of MLPLayers ``in_dim -> dims[0] -> ... -> dims[-1]``, stored as ``sequence``."""

    def __init__(self, in_dim: int, dims: Sequence[int], nonlins: Sequence[Union[str, nn.Module]],
                 batch_norm: bool = True):
        assert len(nonlins) == len(dims)
        super().__init__()
        self.in_dim, self.out_dim, self.dims, self.nonlins = in_dim, dims[-1], dims, nonlins
        widths = [in_dim, *dims]
        self.sequence = nn.Sequential(*[MLPLayer(widths[i], widths[i + 1], nonlins[i], batch_norm)
                                        for i in range(len(dims))])

    def forward(self, x: Tensor) -> Tensor:
        return self.sequence(x.reshape(x.size(0), -1))


class RMLP(nn.Module):
    """x = input(x); x = x + block(x) for every block; return output(x)."""

    def __init__(self, block_in_dim: int, block_dims: Sequence[int], block_nonlins: Sequence[Union[str, nn.Module]],
                 n_blocks: int, out_dim: int, in_dim: int = None, batch_norm: bool = True) -> None:
        super().__init__()
        self.input = nn.Identity() if in_dim is None else MLPLayer(in_dim, block_in_dim, block_nonlins[0], batch_norm)
        self.blocks = nn.ModuleList(MlpBlock(block_in_dim, block_dims, block_nonlins, batch_norm)
                                    for _ in range(n_blocks))
        self.output = nn.Linear(block_dims[-1], out_dim)

    def forward(self, x: Tensor) -> Tensor:
        x = self.input(x)
        for block in self.blocks:
            x = x + block(x)
        return self.output(x)


# --------------------------------------------------------------------------------------------
# loading the reference's checkpoints
# --------------------------------------------------------------------------------------------
_CLASS_MAP = {
    ("wtracker.neural.mlp", "WormPredictor"): WormPredictor,
    ("wtracker.neural.mlp", "MLPLayer"): MLPLayer,
    ("wtracker.neural.mlp", "MlpBlock"): MlpBlock,
    ("wtracker.neural.mlp", "RMLP"): RMLP,
    ("wtracker.neural.config", "IOConfig"): IOConfig,
}


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        hit = _CLASS_MAP.get((module, name))
        return hit if hit is not None else super().find_class(module, name)


class _RefPickle:
    __name__ = "wtracker_b200_ref_pickle"
    Unpickler = _RefUnpickler

    @staticmethod
    def load(f, **kw):
        return _RefUnpickler(f, **kw).load()


def load_worm_predictor(path: str) -> WormPredictor:
    """Loads a whole-module checkpoint written by the reference's Trainer (class paths
    ``wtracker.neural.*``) into the classes above and puts it in eval mode."""
    model = torch.load(path, map_location="cpu", pickle_module=_RefPickle, weights_only=False)
    if not isinstance(model, WormPredictor):
        raise TypeError(f"{path} does not hold a WormPredictor (got {type(model).__name__})")
    model.eval()
    return model
