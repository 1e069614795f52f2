"""ORACLE (test infrastructure, not product code) — plain PyTorch fp32 restatement of the detector
that the reference drives through ``ultralytics.YOLO.predict`` in
``wtracker/sim/sim_controllers/yolo_controller.py:64-90``.

The arithmetic lives in the third-party ``ultralytics`` package (unpinned in the reference's
``pyproject.toml:24`` / ``requirements.yaml:19``; the 8.2.x line is contemporaneous with the
reference's 2024.06.15 release).  It is neither vendored in /root/reference nor installed here, and
the trained ``models/yolov8s_trained.pt`` is stripped from the checkout, so this file restates the
published YOLOv8 algorithm:

  * architecture  : ultralytics/cfg/models/v8/yolov8.yaml (scale "s": depth .33, width .50),
                    nn/modules/conv.py (Conv = Conv2d + BatchNorm(eps=1e-3) + SiLU, fused at load),
                    nn/modules/block.py (C2f, Bottleneck, SPPF), nn/modules/head.py (Detect, DFL)
  * pre-process   : data/augment.py LetterBox (auto=True, stride 32, pad 114) + engine/predictor.py
  * decode        : utils/tal.py make_anchors / dist2bbox
  * post-process  : utils/ops.py non_max_suppression (torchvision.ops.nms), scale_boxes, clip_boxes

PARITY STATUS: *parity unpinned* for the detector — the reference ships no tests, golden vectors or
stored YOLO outputs for this path.  What is pinned: cv2.resize agreement of the
letterbox (tests/test_oracle_cpu.py) and torchvision.ops.nms agreement of the NMS restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# architecture description shared by the oracle model and the weight generator
# --------------------------------------------------------------------------------------------
WIDTH = 0.50
DEPTH = 0.33
REG_MAX = 16


def _ch(c: int) -> int:
    return int(math.ceil(min(c, 1024) * WIDTH / 8) * 8)


def _n(n: int) -> int:
    return max(round(n * DEPTH), 1)


class ConvBnAct(nn.Module):
    """Conv2d(bias=False) + BatchNorm2d(eps=1e-3) + SiLU; ``fuse()`` folds the BN like ultralytics does."""

    def __init__(self, c1: int, c2: int, k: int = 1, s: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.fused = False

    def forward(self, x):
        if self.fused:
            return F.silu(self.conv(x))
        return F.silu(self.bn(self.conv(x)))

    def fuse(self):
        if self.fused:
            return
        w = self.conv.weight.detach()
        bn = self.bn
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        fused = nn.Conv2d(self.conv.in_channels, self.conv.out_channels, self.conv.kernel_size, self.conv.stride,
                          self.conv.padding, bias=True)
        fused.weight.data = (torch.diag(scale) @ w.view(w.shape[0], -1)).view(w.shape)
        fused.bias.data = bn.bias.detach() - bn.weight.detach() * bn.running_mean / torch.sqrt(bn.running_var + bn.eps)
        self.conv = fused
        del self.bn
        self.fused = True


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True):
        super().__init__()
        self.cv1 = ConvBnAct(c1, c2, 3, 1)
        self.cv2 = ConvBnAct(c2, c2, 3, 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = ConvBnAct(c1, 2 * self.c, 1, 1)
        self.cv2 = ConvBnAct((2 + n) * self.c, c2, 1, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = ConvBnAct(c1, c_, 1, 1)
        self.cv2 = ConvBnAct(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        y.extend(self.m(y[-1]) for _ in range(3))
        return self.cv2(torch.cat(y, 1))


class Detect(nn.Module):
    def __init__(self, nc: int, ch: tuple[int, ...]):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = REG_MAX
        self.no = nc + self.reg_max * 4
        c2 = max(16, ch[0] // 4, self.reg_max * 4)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(ConvBnAct(x, c2, 3), ConvBnAct(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(ConvBnAct(x, c3, 3), ConvBnAct(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        self.stride = torch.tensor([8.0, 16.0, 32.0])

    def forward(self, feats):
        return [torch.cat((self.cv2[i](f), self.cv3[i](f)), 1) for i, f in enumerate(feats)]


def make_anchors(shapes, strides):
    """utils/tal.py make_anchors: cell centres, level by level, row-major."""
    pts, strs = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + 0.5
        sy = torch.arange(h, dtype=torch.float32) + 0.5
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        strs.append(torch.full((h * w, 1), float(s)))
    return torch.cat(pts), torch.cat(strs)


def decode_head(levels: list[torch.Tensor], strides=(8, 16, 32), nc: int = 1) -> torch.Tensor:
    """Detect._inference: (B, 64+nc, h, w) x3 -> (B, 4+nc, A) with xywh (centre) in letterboxed pixels."""
    b = levels[0].shape[0]
    no = levels[0].shape[1]
    x_cat = torch.cat([lv.reshape(b, no, -1) for lv in levels], 2)
    box, cls = x_cat.split((REG_MAX * 4, nc), 1)
    anchors, strs = make_anchors([lv.shape[2:] for lv in levels], strides)
    a = box.shape[2]
    prob = box.view(b, 4, REG_MAX, a).transpose(2, 1).softmax(1)               # DFL
    dist = F.conv2d(prob, torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)).view(b, 4, a)
    lt, rb = dist.chunk(2, 1)
    anc = anchors.t().unsqueeze(0)
    x1y1 = anc - lt
    x2y2 = anc + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strs.t().unsqueeze(0)
    return torch.cat((dbox, cls.sigmoid()), 1)


class YoloV8s(nn.Module):
    """DetectionModel for yolov8s with ``nc`` classes; module numbering follows the yaml so that the
    state_dict keys are the ultralytics ones (``model.<i>...``)."""

    def __init__(self, nc: int = 1):
        super().__init__()
        c = [_ch(64), _ch(128), _ch(256), _ch(512), _ch(1024)]   # 32 64 128 256 512
        m = [
            ConvBnAct(3, c[0], 3, 2),                 # 0
            ConvBnAct(c[0], c[1], 3, 2),              # 1
            C2f(c[1], c[1], _n(3), True),             # 2
            ConvBnAct(c[1], c[2], 3, 2),              # 3
            C2f(c[2], c[2], _n(6), True),             # 4
            ConvBnAct(c[2], c[3], 3, 2),              # 5
            C2f(c[3], c[3], _n(6), True),             # 6
            ConvBnAct(c[3], c[4], 3, 2),              # 7
            C2f(c[4], c[4], _n(3), True),             # 8
            SPPF(c[4], c[4], 5),                      # 9
            nn.Upsample(scale_factor=2, mode="nearest"),   # 10
            nn.Identity(),                            # 11 concat [10, 6]
            C2f(c[4] + c[3], c[3], _n(3), False),     # 12
            nn.Upsample(scale_factor=2, mode="nearest"),   # 13
            nn.Identity(),                            # 14 concat [13, 4]
            C2f(c[3] + c[2], c[2], _n(3), False),     # 15
            ConvBnAct(c[2], c[2], 3, 2),              # 16
            nn.Identity(),                            # 17 concat [16, 12]
            C2f(c[2] + c[3], c[3], _n(3), False),     # 18
            ConvBnAct(c[3], c[3], 3, 2),              # 19
            nn.Identity(),                            # 20 concat [19, 9]
            C2f(c[3] + c[4], c[4], _n(3), False),     # 21
            Detect(nc, (c[2], c[3], c[4])),           # 22
        ]
        self.model = nn.ModuleList(m)
        self.nc = nc
        self.widths = c

    def features(self, x, taps: dict | None = None):
        """Returns the three raw head maps (B, 64+nc, h, w); ``taps`` (if given) receives the
        intermediate module outputs by name."""
        m = self.model
        x0 = m[0](x)
        x1 = m[1](x0)
        x2 = m[2](x1)
        x3 = m[3](x2)
        x4 = m[4](x3)
        x = m[5](x4)
        x6 = m[6](x)
        x = m[7](x6)
        x = m[8](x)
        x9 = m[9](x)
        x12 = m[12](torch.cat((m[10](x9), x6), 1))
        x15 = m[15](torch.cat((m[13](x12), x4), 1))
        x18 = m[18](torch.cat((m[16](x15), x12), 1))
        x21 = m[21](torch.cat((m[19](x18), x9), 1))
        if taps is not None:
            taps.update(m0=x0, m1=x1, m2=x2, m3=x3, x4=x4, x6=x6, x9=x9, x12=x12, x15=x15, x18=x18, x21=x21)
        return m[22]([x15, x18, x21])

    def forward(self, x):
        return decode_head(self.features(x), nc=self.nc)

    def fuse(self):
        for mod in self.modules():
            if isinstance(mod, ConvBnAct):
                mod.fuse()
        return self


def synthetic_state_dict(seed: int = 0, nc: int = 1, cls_prior: float = 0.02) -> dict[str, torch.Tensor]:
    """Seeded random YOLOv8s weights in the UNFUSED ultralytics layout (conv + BN statistics),
    rounded to fp16 like an ultralytics checkpoint.  Variance-preserving conv init keeps the
    activations O(1) through the 23 modules; the class-branch bias is set so that roughly
    ``cls_prior`` of the anchors land above conf 0.1."""
    g = torch.Generator().manual_seed(seed)
    model = YoloV8s(nc)
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            out[k] = v.clone()
            continue
        if k.endswith("conv.weight") or (k.endswith(".weight") and v.ndim == 4):
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            t = torch.randn(v.shape, generator=g) * math.sqrt(2.2 / fan_in)
        elif k.endswith("bn.weight"):
            t = 0.8 + 0.4 * torch.rand(v.shape, generator=g)
        elif k.endswith("bn.bias"):
            t = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_mean"):
            t = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            t = 0.8 + 0.4 * torch.rand(v.shape, generator=g)
        elif k.endswith(".bias"):
            if ".cv3." in k:   # class logit bias: logit(0.1) shifted so only the tail crosses conf 0.1
                t = torch.full(v.shape, math.log(0.1 / 0.9) - 1.0)
            else:
                t = 1.0 + 0.1 * torch.randn(v.shape, generator=g)   # box-branch bias (ultralytics inits to 1)
        else:
            t = v.clone()
        out[k] = t.half().float()
    return out


def build_model(state_dict: dict[str, torch.Tensor] | None = None, seed: int = 0, nc: int = 1) -> YoloV8s:
    model = YoloV8s(nc)
    model.load_state_dict(state_dict if state_dict is not None else synthetic_state_dict(seed, nc))
    model.eval()
    model.fuse()
    for p in model.parameters():
        p.requires_grad_(False)
    return model


# --------------------------------------------------------------------------------------------
# pre-process
# --------------------------------------------------------------------------------------------
@dataclass
class LetterboxGeometry:
    src_w: int
    src_h: int
    dst_w: int
    dst_h: int
    new_w: int
    new_h: int
    pad_left: int
    pad_top: int


def letterbox_geometry(src_hw: tuple[int, int], imgsz: int, stride: int = 32, auto: bool = True) -> LetterboxGeometry:
    """LetterBox.__call__ geometry (scaleup=True, center=True)."""
    h, w = src_hw
    r = min(imgsz / h, imgsz / w)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = imgsz - new_w, imgsz - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return LetterboxGeometry(w, h, new_w + left + right, new_h + top + bottom, new_w, new_h, left, top)


def letterbox_cv2(img: np.ndarray, imgsz: int) -> np.ndarray:
    """LetterBox with the real cv2 calls (what ultralytics executes)."""
    import cv2

    g = letterbox_geometry(img.shape[:2], imgsz)
    if (g.new_w, g.new_h) != (g.src_w, g.src_h):
        img = cv2.resize(img, (g.new_w, g.new_h), interpolation=cv2.INTER_LINEAR)
    bottom = g.dst_h - g.new_h - g.pad_top
    right = g.dst_w - g.new_w - g.pad_left
    fill = (114, 114, 114) if img.ndim == 3 else 114
    return cv2.copyMakeBorder(img, g.pad_top, bottom, g.pad_left, right, cv2.BORDER_CONSTANT, value=fill)


def preprocess(frames: list[np.ndarray], imgsz: int) -> torch.Tensor:
    """yolo_controller.py:68-69 (GRAY2BGR) + predictor.preprocess: (n,3,H,W) fp32 in [0,1]."""
    ims = []
    for f in frames:
        if f.ndim == 2:
            f = np.repeat(f[:, :, None], 3, axis=2)   # cv.cvtColor(GRAY2BGR)
        ims.append(letterbox_cv2(np.ascontiguousarray(f), imgsz))
    im = np.stack(ims)[..., ::-1].transpose(0, 3, 1, 2)
    return torch.from_numpy(np.ascontiguousarray(im)).float() / 255


# --------------------------------------------------------------------------------------------
# post-process
# --------------------------------------------------------------------------------------------
def nms_greedy(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float) -> torch.Tensor:
    """Restatement of torchvision.ops.nms: score-descending (lowest index first on ties),
    suppress when IoU > thr (strict).  Returns kept indices into ``boxes``."""
    order = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), i))
    b = boxes.numpy().astype(np.float32)
    areas = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep = []
    for i in order:
        ok = True
        for j in keep:
            xx1, yy1 = max(b[j, 0], b[i, 0]), max(b[j, 1], b[i, 1])
            xx2, yy2 = min(b[j, 2], b[i, 2]), min(b[j, 3], b[i, 3])
            w, h = max(np.float32(0), xx2 - xx1), max(np.float32(0), yy2 - yy1)
            inter = np.float32(w * h)
            if inter / (areas[j] + areas[i] - inter) > np.float32(iou_thres):
                ok = False
                break
        if ok:
            keep.append(i)
    return torch.tensor(keep, dtype=torch.long)


def non_max_suppression(pred: torch.Tensor, conf_thres: float = 0.1, iou_thres: float = 0.7, max_det: int = 1,
                        use_torchvision: bool = True):
    """utils/ops.py non_max_suppression for nc classes, multi_label=False, agnostic=False.
    Returns per image (rows [x1,y1,x2,y2,conf,cls], anchor indices of the rows)."""
    nc = pred.shape[1] - 4
    xc = pred[:, 4:].amax(1) > conf_thres
    pred = pred.transpose(-1, -2).clone()
    xy, wh = pred[..., :2].clone(), pred[..., 2:4].clone()
    pred[..., :2] = xy - wh / 2
    pred[..., 2:4] = xy + wh / 2
    out = []
    for xi in range(pred.shape[0]):
        idx = torch.nonzero(xc[xi]).flatten()
        x = pred[xi][idx]
        if x.shape[0] == 0:
            out.append((torch.zeros((0, 6)), torch.zeros((0,), dtype=torch.long)))
            continue
        conf, j = x[:, 4:4 + nc].max(1, keepdim=True)
        x = torch.cat((x[:, :4], conf, j.float()), 1)
        m = conf.view(-1) > conf_thres
        x, idx = x[m], idx[m]
        if x.shape[0] > 30000:
            o = x[:, 4].argsort(descending=True)[:30000]
            x, idx = x[o], idx[o]
        boxes = x[:, :4] + x[:, 5:6] * 7680
        if use_torchvision:
            import torchvision

            keep = torchvision.ops.nms(boxes, x[:, 4], iou_thres)
        else:
            keep = nms_greedy(boxes, x[:, 4], iou_thres)
        keep = keep[:max_det]
        out.append((x[keep], idx[keep]))
    return out


def scale_boxes(net_hw, boxes: torch.Tensor, img_hw) -> torch.Tensor:
    """utils/ops.py scale_boxes(padding=True) + clip_boxes."""
    gain = min(net_hw[0] / img_hw[0], net_hw[1] / img_hw[1])
    pad_x = round((net_hw[1] - img_hw[1] * gain) / 2 - 0.1)
    pad_y = round((net_hw[0] - img_hw[0] * gain) / 2 - 0.1)
    boxes = boxes.clone()
    boxes[..., 0] -= pad_x
    boxes[..., 2] -= pad_x
    boxes[..., 1] -= pad_y
    boxes[..., 3] -= pad_y
    boxes[..., :4] /= gain
    boxes[..., 0].clamp_(0, img_hw[1])
    boxes[..., 2].clamp_(0, img_hw[1])
    boxes[..., 1].clamp_(0, img_hw[0])
    boxes[..., 3].clamp_(0, img_hw[0])
    return boxes


def scale_params(net_hw, img_hw) -> tuple[float, float, float]:
    gain = min(net_hw[0] / img_hw[0], net_hw[1] / img_hw[1])
    return gain, float(round((net_hw[1] - img_hw[1] * gain) / 2 - 0.1)), float(round((net_hw[0] - img_hw[0] * gain) / 2 - 0.1))


class YoloOracle:
    """The reference's ``YoloController.predict`` (yolo_controller.py:64-90) with the ultralytics
    call replaced by the restatement above."""

    def __init__(self, model: YoloV8s, imgsz: int = 384, conf: float = 0.1, iou: float = 0.7, max_det: int = 1):
        self.model = model
        self.imgsz, self.conf, self.iou, self.max_det = imgsz, conf, iou, max_det

    @torch.no_grad()
    def detect(self, frames: list[np.ndarray]):
        """Per frame: (rows [x1,y1,x2,y2,conf,cls] in camera-view px, anchor indices)."""
        x = preprocess(frames, self.imgsz)
        pred = self.model(x)
        res = non_max_suppression(pred, self.conf, self.iou, self.max_det)
        out = []
        for (rows, idx), f in zip(res, frames):
            rows = rows.clone()
            if rows.shape[0]:
                rows[:, :4] = scale_boxes(x.shape[2:], rows[:, :4], f.shape[:2])
            out.append((rows, idx))
        return out

    def predict(self, frames) -> np.ndarray:
        assert len(frames) > 0
        boxes = []
        for rows, _ in self.detect(list(frames)):
            if rows.shape[0] == 0:
                boxes.append(np.full([4], np.nan))
            else:
                x1, y1, x2, y2 = rows[0, :4].numpy()
                boxes.append(np.array([x1, y1, x2 - x1, y2 - y1], dtype=np.float32))
        return np.stack(boxes, axis=0)
