"""Compiles a YOLOv8 state_dict into the op program the native engine executes
(``wt_buf`` / ``wt_op`` in include/wtracker_b200.h) plus the packed weight blob.

Layout decisions:
  * activations are NHWC bf16; every Concat of the network is a *buffer*: producers store straight
    into their channel slice (TMA store with a channel offset), so concatenation costs nothing;
  * the two nearest-2x upsamples of the neck and the concats after them are not executed: the 1x1 conv that
    consumes concat(upsample(a), b) is evaluated as upsample(conv_a(a)) + conv_b(b) (wt_op.add_buf);
  * C2f: one buffer of (2+n)*c channels holds cv1's two halves and every bottleneck output; the
    bottlenecks read/write channel slices of it, cv2 reads it whole;
  * BatchNorm is folded into the conv weights; weights are bf16 [cout][kh][kw][cin] (K-major rows
    for the UMMA B operand), biases fp32;
  * the first conv sees three identical grey channels scaled by 1/255, so its weights are summed
    over the input channels and divided by 255 on the host and it runs on the u8 image directly;
  * the last 1x1 of the box branch is evaluated in fp32 by the decode kernel, only for anchors that pass
    the confidence filter; the last 1x1
    of the class branch (cout = nc = 1) is a dot product fused into the epilogue of the 3x3 conv
    before it (wt_op.dot_off), which then writes one fp32 logit per anchor.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from wtracker_b200 import _lib as L
from wtracker_b200.detector.arch import REG_MAX, STRIDES, ConvSpec, YoloV8Arch
from wtracker_b200.detector.weights import folded_conv


@dataclass
class Program:
    arch: YoloV8Arch
    net_h: int
    net_w: int
    bufs: list[tuple[int, int, int, int]] = field(default_factory=list)   # (h, w, c, dtype)
    buf_names: list[str] = field(default_factory=list)
    ops: list[dict] = field(default_factory=list)
    blob: bytearray = field(default_factory=bytearray)
    head: list[dict] = field(default_factory=list)    # per level: box buffer, cls feature buffer, cls w/b
    taps: dict[str, tuple[int, int, int]] = field(default_factory=dict)   # module output name -> (buf, coff, c)

    def buf_id(self, name: str) -> int:
        return self.buf_names.index(name)

    @property
    def total_anchors(self) -> int:
        return sum((self.net_h // s) * (self.net_w // s) for s in STRIDES)


def _align(blob: bytearray, a: int) -> int:
    pad = (-len(blob)) % a
    blob.extend(b"\0" * pad)
    return len(blob)


def build_program(sd: dict[str, torch.Tensor], arch: YoloV8Arch, net_h: int, net_w: int, chain: bool = True,
                  chain_exit: bool | None = None) -> Program:
    """``chain``: chain C2f.cv1 onto the stride-2 conv before it where the engine supports it (wt_op.chain_w_off;
    tcgen05 path only, so the scalar validation engine is built with chain=False)."""
    assert net_h % 32 == 0 and net_w % 32 == 0, "network input must be a multiple of 32"
    if chain_exit is None:        # the C2f-exit concat chain follows `chain`
        chain_exit = chain
    if arch.nc != 1:
        raise NotImplementedError(f"the CUDA detector is built for single-class checkpoints (the reference trains with "
                                  f"single_cls: True, yolo_train_config.yaml:27); this one has nc = {arch.nc}")
    p = Program(arch, net_h, net_w)
    specs = {s.name: s for s in arch.conv_specs()}
    c = arch.c

    def new_buf(name: str, down: int, ch: int, dtype: int = L.WT_DT_BF16) -> int:
        p.bufs.append((net_h // down, net_w // down, ch, dtype))
        p.buf_names.append(name)
        return len(p.bufs) - 1

    def add_weights(s: ConvSpec) -> tuple[int, int]:
        w, b = folded_conv(sd, s)
        w_off = _align(p.blob, 16)
        wk = w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)          # [cout][kh][kw][cin]
        p.blob.extend(wk.view(torch.int16).numpy().tobytes())
        b_off = _align(p.blob, 16)
        p.blob.extend(b.float().numpy().tobytes())
        return w_off, b_off

    def add_weights_cat(parts: list[ConvSpec]) -> tuple[int, int]:
        """Several convs over the SAME input stacked along cout (one wider conv)."""
        ws, bs = zip(*(folded_conv(sd, s) for s in parts))
        w_off = _align(p.blob, 16)
        wk = torch.cat(ws, 0).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        p.blob.extend(wk.view(torch.int16).numpy().tobytes())
        b_off = _align(p.blob, 16)
        p.blob.extend(torch.cat(bs, 0).float().numpy().tobytes())
        return w_off, b_off

    def conv_cat(names: list[str], src: tuple[int, int], dst: tuple[int, int]):
        parts = [specs[n] for n in names]
        s0 = parts[0]
        assert all((s.cin, s.k, s.stride, s.bn_act) == (s0.cin, s0.k, s0.stride, s0.bn_act) for s in parts)
        w_off, b_off = add_weights_cat(parts)
        p.ops.append(dict(kind=L.WT_OP_CONV, name="+".join(names), src=src[0], src_coff=src[1], dst=dst[0],
                          dst_coff=dst[1], res=-1, res_coff=0, cin=s0.cin, cout=sum(s.cout for s in parts), k=s0.k,
                          stride=s0.stride, act=L.WT_ACT_SILU if s0.bn_act else L.WT_ACT_NONE, w_off=w_off, b_off=b_off,
                          dot_off=-1))

    def add_weights_slice(s: ConvSpec, c_lo: int, c_hi: int, with_bias: bool) -> tuple[int, int]:
        """Input-channel slice [c_lo, c_hi) of a conv (one term of a conv over a concatenation)."""
        w, b = folded_conv(sd, s)
        w_off = _align(p.blob, 16)
        wk = w[:, c_lo:c_hi].permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        p.blob.extend(wk.view(torch.int16).numpy().tobytes())
        b_off = _align(p.blob, 16)
        p.blob.extend((b if with_bias else torch.zeros_like(b)).float().numpy().tobytes())
        return w_off, b_off

    def conv_over_upsampled_cat(name: str, low: tuple[int, int], c_low: int, high: tuple[int, int], c_high: int,
                                dst: tuple[int, int], down: int):
        """``name`` is a 1x1 conv over concat(upsample2x(low), high).  By linearity it equals
        act(upsample2x(W_low * low + b) + W_high * high): the first term is a 1x1 conv at HALF resolution written
        as f32 (no activation), the second adds it, nearest-upsampled, before its activation (wt_op.add_buf).
        Neither the upsampled tensor nor the concatenation is ever written."""
        s = specs[name]
        assert s.k == 1 and s.stride == 1 and s.cin == c_low + c_high
        part = new_buf(f"{name}.low", down * 2, s.cout, L.WT_DT_F32)
        w_off, b_off = add_weights_slice(s, 0, c_low, with_bias=True)
        p.ops.append(dict(kind=L.WT_OP_CONV, name=name + "[low]", src=low[0], src_coff=low[1], dst=part, dst_coff=0,
                          res=-1, res_coff=0, cin=c_low, cout=s.cout, k=1, stride=1, act=L.WT_ACT_NONE, w_off=w_off,
                          b_off=b_off, dot_off=-1, add_buf=-1, add_coff=0))
        w_off, b_off = add_weights_slice(s, c_low, c_low + c_high, with_bias=False)
        p.ops.append(dict(kind=L.WT_OP_CONV, name=name + "[high]", src=high[0], src_coff=high[1], dst=dst[0],
                          dst_coff=dst[1], res=-1, res_coff=0, cin=c_high, cout=s.cout, k=1, stride=1,
                          act=L.WT_ACT_SILU if s.bn_act else L.WT_ACT_NONE, w_off=w_off, b_off=b_off, dot_off=-1,
                          add_buf=part, add_coff=0))

    def conv(name: str, src: tuple[int, int], dst: tuple[int, int], res: tuple[int, int] | None = None,
             dot: tuple[torch.Tensor, float] | None = None):
        s = specs[name]
        w_off, b_off = add_weights(s)
        dot_off = -1
        if dot is not None:     # fused 1-channel 1x1 head: f32 [cout] weights then the bias
            dot_off = _align(p.blob, 16)
            p.blob.extend(torch.cat([dot[0].reshape(-1).float(), torch.tensor([dot[1]])]).numpy().tobytes())
        p.ops.append(dict(kind=L.WT_OP_CONV, name=name, src=src[0], src_coff=src[1], dst=dst[0], dst_coff=dst[1],
                          res=-1 if res is None else res[0], res_coff=0 if res is None else res[1], cin=s.cin,
                          cout=s.cout, k=s.k, stride=s.stride, act=L.WT_ACT_SILU if s.bn_act else L.WT_ACT_NONE,
                          w_off=w_off, b_off=b_off, dot_off=dot_off))

    chained: list[str] = []

    def c2f(idx: int, src, dst: tuple[int, int], down: int, after: tuple[str, tuple[int, int]] | None = None):
        """``src`` is (buffer, channel offset) or, for the two neck blocks fed by concat(upsample(a), b),
        a dict(low=(buf, coff), c_low=.., high=(buf, coff), c_high=..).  ``after`` = (name, source) of the conv
        whose output is ``src`` and has no other consumer: cv1 is then chained onto it and ``src`` is never written."""
        spec = arch.c2f[idx]
        cc = spec.c
        cat = new_buf(f"c2f{idx}.cat", down, (2 + spec.n) * cc)
        tmp = new_buf(f"c2f{idx}.tmp", down, cc)
        s1 = specs[f"model.{idx}.cv1"]
        if (chain and after is not None and s1.cin == s1.cout and s1.cout in (64, 128) and s1.k == 1
                and specs[after[0]].cout == s1.cin):
            conv(after[0], after[1], (cat, 0))
            w2, b2 = folded_conv(sd, s1)
            w2_off = _align(p.blob, 16)
            p.blob.extend(w2.reshape(s1.cout, s1.cin).contiguous().to(torch.bfloat16).view(torch.int16).numpy().tobytes())
            b2_off = _align(p.blob, 16)
            p.blob.extend(b2.float().numpy().tobytes())
            p.ops[-1].update(name=f"{after[0]}>model.{idx}.cv1", chain_w_off=w2_off, chain_b_off=b2_off,
                             chain_act=L.WT_ACT_SILU if s1.bn_act else L.WT_ACT_NONE)
            chained.append(after[0])
        elif after is not None:
            conv(after[0], after[1], src)
            conv(f"model.{idx}.cv1", src, (cat, 0))
        elif isinstance(src, dict):
            conv_over_upsampled_cat(f"model.{idx}.cv1", src["low"], src["c_low"], src["high"], src["c_high"], (cat, 0),
                                    down)
        else:
            conv(f"model.{idx}.cv1", src, (cat, 0))
        s2 = specs[f"model.{idx}.cv2"]
        cat_chain = (chain_exit and spec.n == 1 and cc == 32 and spec.shortcut and s2.k == 1 and s2.cin == 3 * cc
                     and s2.cout == 2 * cc and net_h % 64 == 0 and net_w % 32 == 0)
        for i in range(spec.n):
            x_in = (cat, (1 + i) * cc)
            conv(f"model.{idx}.m.{i}.cv1", x_in, (tmp, 0))
            if cat_chain:
                # exit of the block as one launch (wt_op.cat_buf): b = y1 + act(conv3x3(tmp)) stays in shared memory and
                # cv2 runs over [y0 | y1 | b] right there; the third slice of the concat buffer is never written
                conv(f"model.{idx}.m.{i}.cv2", (tmp, 0), dst, res=x_in)
                w2, b2 = folded_conv(sd, s2)
                w2_off = _align(p.blob, 16)
                p.blob.extend(w2.reshape(s2.cout, s2.cin).contiguous().to(torch.bfloat16).view(torch.int16).numpy().tobytes())
                b2_off = _align(p.blob, 16)
                p.blob.extend(b2.float().numpy().tobytes())
                p.ops[-1].update(name=f"model.{idx}.m.{i}.cv2>model.{idx}.cv2", chain_w_off=w2_off, chain_b_off=b2_off,
                                 chain_act=L.WT_ACT_SILU if s2.bn_act else L.WT_ACT_NONE, cat_buf=cat, cat_coff=0,
                                 cat_c=2 * cc, chain_cout=s2.cout)
                return
            conv(f"model.{idx}.m.{i}.cv2", (tmp, 0), (cat, (2 + i) * cc), res=x_in if spec.shortcut else None)
        conv(f"model.{idx}.cv2", (cat, 0), dst)

    # ---- buffers that are concat destinations of the neck
    b_in = new_buf("input", 1, 1, L.WT_DT_U8)
    b0 = new_buf("m0", 2, c[0])
    b1 = new_buf("m1", 4, c[1])
    b2 = new_buf("m2", 4, c[1])
    b3 = new_buf("m3", 8, c[2])
    b4 = new_buf("m4", 8, c[2])                   # x4  (concat(up(x12), x4) is never materialised)
    b5 = new_buf("m5", 16, c[3])
    b6 = new_buf("m6", 16, c[3])                  # x6  (concat(up(x9), x6) is never materialised)
    b7 = new_buf("m7", 32, c[4])
    b8 = new_buf("m8", 32, c[4])
    sppf = new_buf("sppf.cat", 32, 2 * c[4])      # [cv1 | pool5 | pool9 | pool13]
    cat20 = new_buf("cat20", 32, c[3] + c[4])     # [x19 | x9]
    cat17 = new_buf("cat17", 16, c[2] + c[3])     # [x16 | x12]
    b15 = new_buf("m15", 8, c[2])
    b18 = new_buf("m18", 16, c[3])
    b21 = new_buf("m21", 32, c[4])

    # ---- backbone
    w0, bias0 = folded_conv(sd, specs["model.0"])
    w_off = _align(p.blob, 16)
    p.blob.extend((w0.sum(1) / 255.0).float().contiguous().numpy().tobytes())      # [cout][3][3]
    b_off = _align(p.blob, 16)
    p.blob.extend(bias0.float().numpy().tobytes())
    # tcgen05 form of the first layer: per output channel [9 taps hi | 0 x 7 | 9 taps lo | 0 x 7] in bf16, w = hi + lo
    w9 = (w0.sum(1) / 255.0).float().reshape(c[0], 9)
    hi = w9.to(torch.bfloat16)
    lo = (w9 - hi.float()).to(torch.bfloat16)
    wmat = torch.zeros((c[0], 32), dtype=torch.bfloat16)
    wmat[:, 0:9] = hi
    wmat[:, 16:25] = lo
    wm_off = _align(p.blob, 16)
    p.blob.extend(wmat.contiguous().view(torch.int16).numpy().tobytes())
    p.ops.append(dict(kind=L.WT_OP_CONV0, name="model.0", src=b_in, src_coff=0, dst=b0, dst_coff=0, res=-1, res_coff=0,
                      cin=1, cout=c[0], k=3, stride=2, act=L.WT_ACT_SILU, w_off=w_off, b_off=b_off,
                      chain_w_off=wm_off if c[0] == 32 else -1))
    # a stride-2 conv of the backbone feeds only the C2f after it: cv1 is chained onto it where the shapes allow
    c2f(2, (b1, 0), (b2, 0), 4, after=("model.1", (b0, 0)))
    c2f(4, (b3, 0), (b4, 0), 8, after=("model.3", (b2, 0)))                    # x4
    c2f(6, (b5, 0), (b6, 0), 16, after=("model.5", (b4, 0)))                   # x6
    c2f(8, (b7, 0), (b8, 0), 32, after=("model.7", (b6, 0)))
    conv("model.9.cv1", (b8, 0), (sppf, 0))
    p.ops.append(dict(kind=L.WT_OP_SPPF_POOL, name="model.9.m", src=sppf, src_coff=0, dst=sppf, dst_coff=c[4] // 2,
                      res=-1, res_coff=0, cin=c[4] // 2, cout=c[4] // 2, k=5, stride=1, act=0, w_off=0, b_off=0))
    conv("model.9.cv2", (sppf, 0), (cat20, c[3]))  # x9

    # ---- neck.  model.10 / model.13 (nearest 2x upsample) and model.11 / model.14 (concat) do not exist as ops:
    # the 1x1 conv that consumes concat(upsample(a), b) is split by linearity (conv_over_upsampled_cat)
    c2f(12, dict(low=(cat20, c[3]), c_low=c[4], high=(b6, 0), c_high=c[3]), (cat17, c[2]), 16)        # x12
    c2f(15, dict(low=(cat17, c[2]), c_low=c[3], high=(b4, 0), c_high=c[2]), (b15, 0), 8)              # x15
    conv("model.16", (b15, 0), (cat17, 0))
    c2f(18, (cat17, 0), (b18, 0), 16)             # x18
    conv("model.19", (b18, 0), (cat20, 0))
    c2f(21, (cat20, 0), (b21, 0), 32)             # x21

    # ---- head
    for lvl, (feat, down) in enumerate(((b15, 8), (b18, 16), (b21, 32))):
        # the first conv of the box branch and of the class branch read the same feature map: ONE conv with
        # box_c + cls_c output channels (a wider N tile re-reads the input tile less often)
        f1 = new_buf(f"head{lvl}.f1", down, arch.box_c + arch.cls_c)
        t1, u1 = f1, f1
        t2 = new_buf(f"head{lvl}.box2", down, arch.box_c)
        logit = new_buf(f"head{lvl}.cls", down, 1, L.WT_DT_F32)
        conv_cat([f"model.22.cv2.{lvl}.0", f"model.22.cv3.{lvl}.0"], (feat, 0), (f1, 0))
        conv(f"model.22.cv2.{lvl}.1", (t1, 0), (t2, 0))
        # the last 1x1 of the box branch (64 DFL logits per anchor) is NOT run over the map: the decode kernel
        # evaluates it for the few anchors that pass the confidence filter (wt_head_level.box_feat)
        wb, bb = folded_conv(sd, specs[f"model.22.cv2.{lvl}.2"])
        bw_off = _align(p.blob, 16)
        p.blob.extend(wb.reshape(wb.shape[0], -1).to(torch.bfloat16).contiguous().view(torch.int16).numpy().tobytes())
        bb_off = _align(p.blob, 16)
        p.blob.extend(bb.float().numpy().tobytes())
        # the class branch ends in a 1x1 conv with nc = 1 output: a dot product fused into the epilogue of
        # the conv before it (fp32 weights on the fp32 accumulator, the 128-channel feature map is never stored)
        wc, bc = folded_conv(sd, specs[f"model.22.cv3.{lvl}.2"])
        conv(f"model.22.cv3.{lvl}.1", (u1, arch.box_c), (logit, 0), dot=(wc, float(bc.reshape(-1)[0])))
        p.head.append(dict(box_feat=t2, box_w_off=bw_off, box_b_off=bb_off, box_c=arch.box_c, cls_logit=logit,
                           h=net_h // down, w=net_w // down, stride=STRIDES[lvl]))
    _align(p.blob, 16)
    _assign_lanes(p)

    p.taps = {
        "x4": (b4, 0, c[2]), "x6": (b6, 0, c[3]), "x9": (cat20, c[3], c[4]),
        "x12": (cat17, c[2], c[3]), "x15": (b15, 0, c[2]), "x18": (b18, 0, c[3]), "x21": (b21, 0, c[4]),
        "m0": (b0, 0, c[0]), "m1": (b1, 0, c[1]), "m2": (b2, 0, c[1]), "m3": (b3, 0, c[2]),
    }
    for name, tap in (("model.1", "m1"), ("model.3", "m3")):
        if name in chained:
            del p.taps[tap]      # never written: it only exists as tiles inside the chained kernel
    return p


def _assign_lanes(p: Program) -> None:
    """Two launch lanes (wt_op.lane) for the part of the graph that has independent branches: once x15 exists, the
    heads of pyramid levels 0 and 1 (lane 1, the engine's side stream) run beside the rest of the neck and the
    level-2 head (lane 0).  Every kernel is persistent with one CTA per SM, so two lanes do not share SMs; what
    overlaps is the drain of one kernel (last tiles, partial last wave) with the ramp-up of the other lane's next
    kernel, which on one stream is a dependency bubble after every launch.  The order stays topological (the engine
    derives cross-lane waits from earlier ops only) and alternates lanes so both queues stay fed."""
    names = [o["name"] for o in p.ops]
    if "model.16" not in names:
        return
    start = names.index("model.16")
    tail = p.ops[start:]
    head = {lvl: [o for o in tail if o["name"].startswith(("model.22.cv2.%d." % lvl, "model.22.cv3.%d." % lvl))]
            for lvl in range(3)}
    neck = [o for o in tail if not o["name"].startswith("model.22.")]
    x18 = max(i for i, o in enumerate(neck) if o["name"].startswith("model.18."))   # model.18.cv2 writes x18
    for o in head[0] + head[1]:
        o["lane"] = 1
    order: list[dict] = []
    side = list(head[0])
    for i, o in enumerate(neck):
        order.append(o)
        if i == x18:
            side += head[1]          # level 1 reads x18: only after its producer in program order
        if side:
            order.append(side.pop(0))
    order += side + head[2]
    assert len(order) == len(tail) and {id(o) for o in order} == {id(o) for o in tail}
    p.ops[start:] = order


def ops_as_ctypes(p: Program):
    bufs = (L.WtBuf * len(p.bufs))(*[L.WtBuf(*b) for b in p.bufs])
    ops = (L.WtOp * len(p.ops))()
    for i, o in enumerate(p.ops):
        ops[i] = L.WtOp(o["kind"], o["src"], o["src_coff"], o["dst"], o["dst_coff"], o["res"], o["res_coff"],
                        o["cin"], o["cout"], o["k"], o["stride"], o["act"], o["w_off"], o["b_off"], o.get("dot_off", -1), o.get("add_buf", -1), o.get("add_coff", 0),
                        o.get("lane", 0), o.get("chain_act", 0), o.get("chain_w_off", -1), o.get("chain_b_off", -1),
                        o.get("cat_buf", -1), o.get("cat_coff", 0), o.get("cat_c", 0), o.get("chain_cout", 0))
    return bufs, ops


def blob_tensor(p: Program) -> torch.Tensor:
    return torch.from_numpy(np.frombuffer(bytes(p.blob), dtype=np.uint8).copy())
