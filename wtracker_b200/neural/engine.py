"""CUDA execution of a WormPredictor (K9): folds every BatchNorm1d into its Linear, packs the
layers in execution order and calls ``wt_resmlp_forward`` (one sample per thread, weights in shared
memory).  Replaces ``model.forward(Tensor(boxes))`` of MLPController.provide_movement_vector
(reference: wtracker/sim/sim_controllers/mlp_controllers.py:59)."""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from wtracker_b200 import _lib as L
from wtracker_b200.neural.mlp import MLPLayer, MlpBlock, RMLP, WormPredictor


def _fold(layer: MLPLayer) -> tuple[torch.Tensor, torch.Tensor, bool]:
    """(W [out, in], b [out], relu?) with the BatchNorm folded (eval-mode statistics)."""
    lin = layer.mlp_layer[0]
    w, b = lin.weight.detach().double(), lin.bias.detach().double()
    relu = False
    for m in list(layer.mlp_layer)[1:]:
        if isinstance(m, nn.BatchNorm1d):
            s = m.weight.detach().double() / torch.sqrt(m.running_var.detach().double() + m.eps)
            w = w * s[:, None]
            b = (b - m.running_mean.detach().double()) * s + m.bias.detach().double()
        elif isinstance(m, nn.ReLU):
            relu = True
        elif not isinstance(m, nn.Identity):
            raise NotImplementedError(f"activation {type(m).__name__} is not supported by the CUDA ResMLP")
    return w.float(), b.float(), relu


def pack_resmlp(model: WormPredictor) -> tuple[np.ndarray, dict]:
    """Flat fp32 weight blob + shape description for ``wt_resmlp_desc``."""
    net = model.model
    if not isinstance(net, RMLP) or not isinstance(net.input, MLPLayer):
        raise NotImplementedError("the CUDA predictor expects an RMLP with an input layer")
    chunks: list[torch.Tensor] = []
    w, b, relu = _fold(net.input)
    assert relu, "input layer must end in ReLU"
    in_dim, hidden = w.shape[1], w.shape[0]
    chunks += [w.reshape(-1), b]
    block_dims: list[int] = []
    for bi, block in enumerate(net.blocks):
        assert isinstance(block, MlpBlock)
        dims = []
        for layer in block.sequence:
            w, b, relu = _fold(layer)
            assert relu, "block layers must end in ReLU"
            dims.append(w.shape[0])
            chunks += [w.reshape(-1), b]
        if bi == 0:
            block_dims = dims
        assert dims == block_dims and dims[-1] == hidden
    chunks += [net.output.weight.detach().float().reshape(-1), net.output.bias.detach().float()]
    blob = torch.cat(chunks).contiguous().numpy()
    desc = dict(in_dim=in_dim, hidden=hidden, out_dim=net.output.out_features, n_blocks=len(net.blocks),
                block_len=len(block_dims), block_dims=block_dims)
    return blob, desc


def transpose_blob(blob: np.ndarray, d: dict) -> np.ndarray:
    """The same layers with every weight matrix stored [in][out] (then bias[out]), zero-padded to a multiple of 4
    floats: the layout ``wt_hot_tail`` stages with plain 16-byte copies (wt_tail_args.weights_t)."""
    dims = [d["hidden"]] + list(d["block_dims"]) * d["n_blocks"] + [d["out_dim"]]
    out, off, nin = [], 0, d["in_dim"]
    for nout in dims:
        w = blob[off: off + nout * nin].reshape(nout, nin)
        out += [np.ascontiguousarray(w.T).reshape(-1), blob[off + nout * nin: off + nout * nin + nout]]
        off += nout * nin + nout
        nin = nout
    assert off == blob.size
    t = np.concatenate(out).astype(np.float32)
    return np.concatenate([t, np.zeros((-t.size) % 4, np.float32)])


class ResMLPEngine:
    def __init__(self, model: WormPredictor, device: str = "cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("wtracker_b200 needs a CUDA device (no CPU fallback)")
        self.lib = L.lib()
        self.device = torch.device(device)
        self.io_config = model.io_config
        blob, d = pack_resmlp(model)
        self.shape = d
        self.weights = torch.from_numpy(blob).to(self.device)
        self.weights_t = torch.from_numpy(transpose_blob(blob, d)).to(self.device)
        dims = (C.c_int32 * 8)(*(d["block_dims"] + [0] * (8 - len(d["block_dims"]))))
        self._desc = L.WtResmlpDesc(d["in_dim"], d["hidden"], d["out_dim"], d["n_blocks"], d["block_len"], dims,
                                    self.weights.data_ptr(), int(self.weights.numel()))
        self.desc = self._desc
        self._host_in = torch.empty((1, d["in_dim"]), dtype=torch.float32).pin_memory()
        self._host_out = torch.empty((1, d["out_dim"]), dtype=torch.float32).pin_memory()
        self._dev_in = torch.empty((1, d["in_dim"]), dtype=torch.float32, device=self.device)
        self._dev_out = torch.empty((1, d["out_dim"]), dtype=torch.float32, device=self.device)

    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """x: cuda fp32 [n, in_dim] -> cuda fp32 [n, out_dim]."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == self.shape["in_dim"]
        n = x.shape[0]
        if out is None:
            out = torch.empty((n, self.shape["out_dim"]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(self.lib.wt_resmlp_forward(C.byref(self._desc), x.data_ptr(), out.data_ptr(), n,
                                               torch.cuda.current_stream().cuda_stream), "wt_resmlp_forward")
        return out

    def forward_host(self, x: np.ndarray) -> np.ndarray:
        """Single-row convenience path used by MLPController (one prediction per cycle)."""
        x = np.asarray(x, dtype=np.float32).reshape(-1, self.shape["in_dim"])
        if x.shape[0] != 1:
            dev = torch.from_numpy(np.ascontiguousarray(x)).to(self.device)
            return self.forward(dev).cpu().numpy()
        with torch.cuda.device(self.device):
            self._host_in.copy_(torch.from_numpy(x))
            self._dev_in.copy_(self._host_in, non_blocking=True)
            self.forward(self._dev_in, self._dev_out)
            self._host_out.copy_(self._dev_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return self._host_out.numpy().copy()
