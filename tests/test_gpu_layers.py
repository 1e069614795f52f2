"""K5 launch by launch: EVERY op of the compiled detector program — plain, residual, upsampled-addend, dot-head, chained
and concat-chained convolutions, layer 0 on the tensor cores, the SPPF pool — against a plain PyTorch fp32 statement of
the same op applied to the op's OWN input (the bf16 / u8 / f32 buffers the kernel read), so that an error cannot hide
behind, or be blamed on, the layers before it.  The program is stepped op by op (`wt_engine_forward(first, last)`), because
some buffers are reused later in the pass.

Tolerance per op (printed, and written to gpurun_out/layer_parity.json): the kernel accumulates in fp32 from the same bf16
operands; what differs is the accumulation order, tanh.approx in SiLU and the final rounding to bf16 (2^-9 relative), so
mean |err| <= 0.3 % of the reference's standard deviation and max |err| <= 1.5 % of its largest magnitude (measured
worst over the 54 ops at both network sizes: 0.126 % and 0.53 %).  A wrong tap, a swapped channel
block or a missing residual gives errors of order 100 %."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_common import synthetic_sd, views_for

pytestmark = pytest.mark.gpu

MEAN_TOL, MAX_TOL = 3e-3, 1.5e-2


def blob_f32(blob, off, n):
    return torch.from_numpy(blob[off: off + 4 * n].view(np.float32).copy())


def blob_bf16(blob, off, shape):
    n = int(np.prod(shape))
    return torch.from_numpy(blob[off: off + 2 * n].view(np.int16).copy()).view(torch.bfloat16).float().view(*shape)


def act(x, kind):
    return F.silu(x) if kind == 1 else x


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("view,imgsz", [(640, 640), (360, 384)])
def test_every_launch_against_torch_fp32_on_its_own_input(view, imgsz):
    from wtracker_b200 import _lib as L
    from wtracker_b200.detector.engine import DetectorEngine

    n = 3
    eng = DetectorEngine(synthetic_sd(), (view, view), imgsz, batch=n, max_det=1)
    eng.detect_views(views_for(view, n))                 # fills the input buffer (and warms every kernel up)
    blob = eng.weights.cpu().numpy()
    prog = eng.program
    report, failures = [], []
    for i, o in enumerate(prog.ops):
        eng.forward(n, i, i + 1)
        torch.cuda.synchronize()
        src = eng.buffer_tensor(o["src"], n)
        dst = eng.buffer_tensor(o["dst"], n)
        name, cin, cout, k, s = o["name"], o["cin"], o["cout"], o["k"], o["stride"]
        if o["kind"] == L.WT_OP_SPPF_POOL:
            x = nchw(src[..., o["src_coff"]: o["src_coff"] + cin])
            ref = []
            for _ in range(3):
                x = F.max_pool2d(x, 5, 1, 2)
                ref.append(x)
            ref = torch.cat(ref, 1)
            got = nchw(dst[..., o["dst_coff"]: o["dst_coff"] + 3 * cin])
        elif o["kind"] == L.WT_OP_CONV0:
            # grey u8 input; weights = f32 [cout][3][3] with GRAY2BGR and /255 folded in (program.py), bias f32
            w = blob_f32(blob, o["w_off"], cout * 9).view(cout, 1, 3, 3)
            b = blob_f32(blob, o["b_off"], cout)
            ref = act(F.conv2d(src.float().permute(0, 3, 1, 2).cpu(), w, b, stride=2, padding=1), o["act"])
            got = nchw(dst[..., o["dst_coff"]: o["dst_coff"] + cout]).cpu()
        else:
            x = nchw(src[..., o["src_coff"]: o["src_coff"] + cin]).cpu()
            w = blob_bf16(blob, o["w_off"], (cout, k, k, cin)).permute(0, 3, 1, 2).contiguous()
            b = blob_f32(blob, o["b_off"], cout)
            y = F.conv2d(x, w, b, stride=s, padding=k // 2)
            if o.get("add_buf", -1) >= 0:                 # + nearest-2x upsampled f32 partial sums, before the activation
                add = nchw(eng.buffer_tensor(o["add_buf"], n)[..., o["add_coff"]: o["add_coff"] + cout]).cpu()
                y = y + F.interpolate(add, scale_factor=2, mode="nearest")
            y = act(y, o["act"])
            if o["res"] >= 0:                             # residual AFTER the activation (Bottleneck: x + cv2(cv1(x)))
                res = eng.buffer_tensor(o["res"], n) if o["res"] != o["dst"] else dst
                y = y + nchw(res[..., o["res_coff"]: o["res_coff"] + cout]).cpu()
            if o.get("dot_off", -1) >= 0:                 # class head: 1-channel 1x1 conv on the fp32 activations
                wd = blob_f32(blob, o["dot_off"], cout + 1)
                ref = (y * wd[:cout].view(1, -1, 1, 1)).sum(1, keepdim=True) + wd[cout]
                got = nchw(dst[..., :1]).cpu()
            elif o.get("chain_w_off", -1) >= 0:
                y = y.to(torch.bfloat16).float()          # the intermediate map is rounded exactly as if it had been stored
                if o.get("cat_buf", -1) >= 0:             # concat chain: 1x1 conv over [cat slice | y]
                    cat = nchw(eng.buffer_tensor(o["cat_buf"], n)[..., o["cat_coff"]: o["cat_coff"] + o["cat_c"]]).cpu()
                    y = torch.cat((cat, y), 1)
                    c2 = o["chain_cout"]
                else:
                    c2 = cout
                w2 = blob_bf16(blob, o["chain_w_off"], (c2, y.shape[1])).view(c2, y.shape[1], 1, 1)
                ref = act(F.conv2d(y, w2, blob_f32(blob, o["chain_b_off"], c2)), o["chain_act"])
                got = nchw(dst[..., o["dst_coff"]: o["dst_coff"] + c2]).cpu()
            else:
                ref = y
                got = nchw(dst[..., o["dst_coff"]: o["dst_coff"] + cout]).cpu()
        ref, got = ref.cpu(), got.cpu()
        assert ref.shape == got.shape, (name, ref.shape, got.shape)
        err = (got - ref).abs()
        mean_rel = float(err.mean() / ref.std().clamp_min(1e-6))
        max_rel = float(err.max() / ref.abs().max().clamp_min(1e-6))
        report.append(dict(op=i, name=name, mean_err_over_std=round(mean_rel, 5), max_err_over_max=round(max_rel, 5)))
        if not (mean_rel <= MEAN_TOL and max_rel <= MAX_TOL):
            failures.append(report[-1])
    worst = max(report, key=lambda r: r["max_err_over_max"])
    print(f"[layer parity {view}->{imgsz}] {len(report)} ops, worst {worst}")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", f"layer_parity_{imgsz}.json"), "w") as f:
        json.dump(report, f, indent=0)
    assert len(report) == len(prog.ops) >= 50
    assert not failures, failures[:6]
