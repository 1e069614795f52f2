"""YOLOv8 detection architecture as data: which convolutions exist, their ultralytics state_dict
names and shapes.  (ultralytics/cfg/models/v8/yolov8.yaml + nn/modules/{conv,block,head}.py — the
network the reference runs through ``YOLO.predict`` in
wtracker/sim/sim_controllers/yolo_controller.py:72-78.)

Only the description lives here; the CUDA program is built from it in ``program.py``.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

SCALES = {  # depth, width, max_channels
    "n": (0.33, 0.25, 1024),
    "s": (0.33, 0.50, 1024),
    "m": (0.67, 0.75, 768),
    "l": (1.00, 1.00, 512),
    "x": (1.00, 1.25, 512),
}
REG_MAX = 16
STRIDES = (8, 16, 32)


@dataclass(frozen=True)
class ConvSpec:
    name: str        # state_dict prefix, e.g. "model.2.m.0.cv1" (=> .conv.weight / .bn.*) or "model.22.cv2.0.2"
    cin: int
    cout: int
    k: int
    stride: int
    bn_act: bool     # True: Conv2d(no bias)+BN+SiLU; False: plain Conv2d with bias (head outputs)


@dataclass(frozen=True)
class C2fSpec:
    idx: int
    c1: int
    c2: int
    n: int
    shortcut: bool

    @property
    def c(self) -> int:
        return self.c2 // 2


class YoloV8Arch:
    """Channel plan of yolov8<scale> with ``nc`` classes."""

    def __init__(self, scale: str = "s", nc: int = 1):
        depth, width, max_ch = SCALES[scale]
        self.scale, self.nc = scale, nc

        def ch(c: int) -> int:
            return int(math.ceil(min(c, max_ch) * width / 8) * 8)

        def rep(n: int) -> int:
            return max(round(n * depth), 1)

        self.c = [ch(64), ch(128), ch(256), ch(512), ch(1024)]
        c = self.c
        self.c2f = {
            2: C2fSpec(2, c[1], c[1], rep(3), True),
            4: C2fSpec(4, c[2], c[2], rep(6), True),
            6: C2fSpec(6, c[3], c[3], rep(6), True),
            8: C2fSpec(8, c[4], c[4], rep(3), True),
            12: C2fSpec(12, c[4] + c[3], c[3], rep(3), False),
            15: C2fSpec(15, c[3] + c[2], c[2], rep(3), False),
            18: C2fSpec(18, c[2] + c[3], c[3], rep(3), False),
            21: C2fSpec(21, c[3] + c[4], c[4], rep(3), False),
        }
        self.head_ch = (c[2], c[3], c[4])
        self.box_c = max(16, self.head_ch[0] // 4, REG_MAX * 4)
        self.cls_c = max(self.head_ch[0], min(nc, 100))

    # ------------------------------------------------------------------ enumeration
    def conv_specs(self) -> list[ConvSpec]:
        """Every convolution, in module order."""
        c = self.c
        out: list[ConvSpec] = []

        def conv(name, cin, cout, k, s):
            out.append(ConvSpec(name, cin, cout, k, s, True))

        def c2f(spec: C2fSpec):
            p = f"model.{spec.idx}"
            conv(f"{p}.cv1", spec.c1, 2 * spec.c, 1, 1)
            conv(f"{p}.cv2", (2 + spec.n) * spec.c, spec.c2, 1, 1)
            for i in range(spec.n):
                conv(f"{p}.m.{i}.cv1", spec.c, spec.c, 3, 1)
                conv(f"{p}.m.{i}.cv2", spec.c, spec.c, 3, 1)

        conv("model.0", 3, c[0], 3, 2)
        conv("model.1", c[0], c[1], 3, 2)
        c2f(self.c2f[2])
        conv("model.3", c[1], c[2], 3, 2)
        c2f(self.c2f[4])
        conv("model.5", c[2], c[3], 3, 2)
        c2f(self.c2f[6])
        conv("model.7", c[3], c[4], 3, 2)
        c2f(self.c2f[8])
        conv("model.9.cv1", c[4], c[4] // 2, 1, 1)
        conv("model.9.cv2", c[4] * 2, c[4], 1, 1)
        c2f(self.c2f[12])
        c2f(self.c2f[15])
        conv("model.16", c[2], c[2], 3, 2)
        c2f(self.c2f[18])
        conv("model.19", c[3], c[3], 3, 2)
        c2f(self.c2f[21])
        for lvl, x in enumerate(self.head_ch):
            conv(f"model.22.cv2.{lvl}.0", x, self.box_c, 3, 1)
            conv(f"model.22.cv2.{lvl}.1", self.box_c, self.box_c, 3, 1)
            out.append(ConvSpec(f"model.22.cv2.{lvl}.2", self.box_c, 4 * REG_MAX, 1, 1, False))
            conv(f"model.22.cv3.{lvl}.0", x, self.cls_c, 3, 1)
            conv(f"model.22.cv3.{lvl}.1", self.cls_c, self.cls_c, 3, 1)
            out.append(ConvSpec(f"model.22.cv3.{lvl}.2", self.cls_c, self.nc, 1, 1, False))
        return out

    def macs_per_image(self, net_h: int, net_w: int) -> int:
        """Multiply-accumulates of all convolutions for one (net_h, net_w) image."""
        res = {}
        h, w = net_h, net_w
        total = 0
        # spatial size of each conv's OUTPUT follows from the module it sits in
        def out_hw(name: str) -> tuple[int, int]:
            i = int(name.split(".")[1])
            down = {0: 2, 1: 4, 2: 4, 3: 8, 4: 8, 5: 16, 6: 16, 7: 32, 8: 32, 9: 32, 12: 16, 15: 8, 16: 16, 18: 16,
                    19: 32, 21: 32}
            if i == 22:
                lvl = int(name.split(".")[3])
                d = STRIDES[lvl]
            else:
                d = down[i]
            return net_h // d, net_w // d

        for s in self.conv_specs():
            oh, ow = out_hw(s.name)
            total += oh * ow * s.cout * s.cin * s.k * s.k
        return total
