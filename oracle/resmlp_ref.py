"""ORACLE (test infrastructure, not product code) — numpy fp32 restatement of the ResMLP forward
pass: WormPredictor.forward -> RMLP.forward (/root/reference/wtracker/neural/mlp.py:47-48,176-188),
MlpBlock (:92-141), MLPLayer = Linear -> BatchNorm1d(eval) -> ReLU (:51-89).  BatchNorm is applied
as a separate step (NOT folded) so that this checks the product's folding too.
Pinned by the known-answer vectors in tests/golden/reference_golden.npz (made by the real reference).
"""

from __future__ import annotations

import numpy as np


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def _layer(seq, x: np.ndarray) -> np.ndarray:
    """One MLPLayer's nn.Sequential: Linear, optional BatchNorm1d (running statistics), activation."""
    import torch.nn as nn

    for m in seq:
        if isinstance(m, nn.Linear):
            x = x @ _np(m.weight).T + _np(m.bias)
        elif isinstance(m, nn.BatchNorm1d):
            x = (x - _np(m.running_mean)) / np.sqrt(_np(m.running_var) + np.float32(m.eps)) * _np(m.weight) + _np(m.bias)
        elif isinstance(m, nn.ReLU):
            x = np.maximum(x, np.float32(0))
        elif isinstance(m, nn.Identity):
            pass
        else:
            raise NotImplementedError(type(m).__name__)
    return x.astype(np.float32)


def resmlp_forward(predictor, x: np.ndarray) -> np.ndarray:
    """predictor: a loaded WormPredictor module (used only as a weight container)."""
    net = predictor.model
    x = np.asarray(x, dtype=np.float32)
    x = _layer(net.input.mlp_layer, x)
    for block in net.blocks:
        y = x
        for layer in block.sequence:
            y = _layer(layer.mlp_layer, y)
        x = x + y
    return x @ _np(net.output.weight).T + _np(net.output.bias)
