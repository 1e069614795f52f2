"""Detector weights: loading an ultralytics checkpoint without ultralytics, seeded synthetic
weights (the trained ``models/yolov8s_trained.pt`` is not in the reference checkout), and BatchNorm
folding (what ultralytics does at load time with ``model.fuse()``).
"""

from __future__ import annotations

import io
import json
import math
import pickle
from pathlib import Path

import torch

from wtracker_b200.detector.arch import ConvSpec, YoloV8Arch

CALIB_PATH = Path(__file__).resolve().parents[2] / "models" / "synthetic_yolov8s_calib.json"


# --------------------------------------------------------------------------------------------
# synthetic weights
# --------------------------------------------------------------------------------------------
def synthetic_state_dict(seed: int = 0, arch: YoloV8Arch | None = None, calibrated: bool = True,
                         head_gain: float = 1.0) -> dict[str, torch.Tensor]:
    """Seeded random weights in the UNFUSED ultralytics layout (conv weight + BN statistics),
    rounded through fp16 like an ultralytics checkpoint.  Conv weights ~ N(0, 2.2/fan_in) keep the
    activations O(1) through all modules.  With ``calibrated`` the final 1x1 head convolutions are
    replaced by the ones in ``models/synthetic_yolov8s_calib.json`` (made by
    ``tests/tools/calibrate_synthetic.py``: principal directions of the head features) so that ~1-3 % of
    the anchors of a synthetic worm frame pass conf 0.1 and the DFL distributions are not flat."""
    arch = arch or YoloV8Arch("s", 1)
    g = torch.Generator().manual_seed(seed)
    sd: dict[str, torch.Tensor] = {}

    def rnd(*shape):
        return torch.randn(*shape, generator=g)

    def uni(*shape):
        return torch.rand(*shape, generator=g)

    for s in arch.conv_specs():
        fan_in = s.cin * s.k * s.k
        w = rnd(s.cout, s.cin, s.k, s.k) * math.sqrt(2.2 / fan_in)
        if s.bn_act:
            sd[f"{s.name}.conv.weight"] = w
            sd[f"{s.name}.bn.weight"] = 0.8 + 0.4 * uni(s.cout)
            sd[f"{s.name}.bn.bias"] = 0.1 * rnd(s.cout)
            sd[f"{s.name}.bn.running_mean"] = 0.1 * rnd(s.cout)
            sd[f"{s.name}.bn.running_var"] = 0.8 + 0.4 * uni(s.cout)
        else:
            sd[f"{s.name}.weight"] = w
            if ".cv3." in s.name:
                sd[f"{s.name}.bias"] = torch.full((s.cout,), math.log(0.1 / 0.9) - 1.0)
            else:
                sd[f"{s.name}.bias"] = 1.0 + 0.1 * rnd(s.cout)
    if not calibrated and head_gain != 1.0:
        # purely random heads answer with nearly constant class logits (std ~0.03 over an image); the gain spreads them
        # so that a share of the anchors crosses conf 0.1 — a detector whose every decision is a near-tie (parity tests)
        for lvl in range(3):
            sd[f"model.22.cv3.{lvl}.2.weight"] = sd[f"model.22.cv3.{lvl}.2.weight"] * head_gain
    if calibrated and CALIB_PATH.exists():
        calib = json.loads(CALIB_PATH.read_text()).get(f"{arch.scale}-nc{arch.nc}-seed{seed}")
        if calib:
            for name, repl in calib.items():
                w = torch.tensor(repl["weight"], dtype=torch.float32)
                sd[f"{name}.weight"] = w.view(w.shape[0], w.shape[1], 1, 1)
                sd[f"{name}.bias"] = torch.tensor(repl["bias"], dtype=torch.float32)
    return {k: v.half().float() for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
# ultralytics checkpoints without ultralytics
# --------------------------------------------------------------------------------------------
class _Shim(torch.nn.Module):
    """Stand-in for any ultralytics class found in a checkpoint pickle; keeps sub-modules and
    tensors so that ``state_dict()`` works, ignores everything else."""

    def __init__(self, *a, **k):
        super().__init__()

    def __setstate__(self, state):
        self.__dict__.update(state)
        for k in ("_parameters", "_buffers", "_modules"):
            self.__dict__.setdefault(k, {})


class _ShimUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if module.split(".")[0] == "ultralytics":
            return type(name, (_Shim,), {"__module__": module})
        return super().find_class(module, name)


class _ShimPickleModule:
    __name__ = "wtracker_b200_shim_pickle"
    Unpickler = _ShimUnpickler

    @staticmethod
    def load(f, **kw):
        return _ShimUnpickler(f, **kw).load()


def load_ultralytics_checkpoint(path: str) -> dict[str, torch.Tensor]:
    """Returns the fp32 state_dict of the DetectionModel stored in an ultralytics ``.pt`` file
    (a pickled dict whose "model" (or "ema") entry is the whole module, usually in fp16)."""
    ckpt = torch.load(path, map_location="cpu", pickle_module=_ShimPickleModule, weights_only=False)
    model = ckpt.get("ema") or ckpt["model"] if isinstance(ckpt, dict) else ckpt
    sd = model.state_dict() if hasattr(model, "state_dict") else model
    return {k: v.float() for k, v in sd.items() if torch.is_tensor(v)}


def infer_arch(sd: dict[str, torch.Tensor]) -> YoloV8Arch:
    """Recovers scale and class count from tensor shapes."""
    c0 = sd["model.0.conv.weight"].shape[0]
    nc = sd["model.22.cv3.0.2.weight"].shape[0]
    for scale in ("n", "s", "m", "l", "x"):
        a = YoloV8Arch(scale, nc)
        if a.c[0] == c0 and all(
            tuple(sd[f"{s.name}.conv.weight" if s.bn_act else f"{s.name}.weight"].shape) == (s.cout, s.cin, s.k, s.k)
            for s in a.conv_specs()
        ):
            return a
    raise ValueError("state_dict does not match any YOLOv8 detection scale")


# --------------------------------------------------------------------------------------------
# BatchNorm folding
# --------------------------------------------------------------------------------------------
BN_EPS = 1e-3


def folded_conv(sd: dict[str, torch.Tensor], s: ConvSpec) -> tuple[torch.Tensor, torch.Tensor]:
    """(weight [cout, cin, k, k], bias [cout]) in fp32 with the BatchNorm folded in."""
    if not s.bn_act:
        return sd[f"{s.name}.weight"].float(), sd[f"{s.name}.bias"].float()
    w = sd[f"{s.name}.conv.weight"].float()
    if f"{s.name}.bn.weight" not in sd:   # already fused checkpoint
        return w, sd[f"{s.name}.conv.bias"].float()
    gamma, beta = sd[f"{s.name}.bn.weight"].float(), sd[f"{s.name}.bn.bias"].float()
    mean, var = sd[f"{s.name}.bn.running_mean"].float(), sd[f"{s.name}.bn.running_var"].float()
    scale = gamma / torch.sqrt(var + BN_EPS)
    return w * scale.view(-1, 1, 1, 1), beta - mean * scale


def save_state_dict(sd: dict[str, torch.Tensor], path: str) -> None:
    buf = io.BytesIO()
    torch.save(sd, buf)
    Path(path).write_bytes(buf.getvalue())
