"""K6-K8 parity: the decode + NMS kernel fed with the ORACLE's fp32 head tensors must keep exactly
the anchors torchvision.ops.nms keeps (bit-exact indices, order included)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import yolov8_ref as Y

pytestmark = pytest.mark.gpu


def random_head(n, hw_levels, seed, logit_mean=-3.0, logit_std=1.2, box_std=2.0):
    g = torch.Generator().manual_seed(seed)
    levels = []
    for h, w in hw_levels:
        box = torch.randn(n, 64, h, w, generator=g) * box_std
        cls = torch.randn(n, 1, h, w, generator=g) * logit_std + logit_mean
        levels.append(torch.cat([box, cls], 1))
    return levels


def run_post(levels, net_hw, img_hw, conf, iou, max_det):
    from wtracker_b200 import _lib as L
    from wtracker_b200.detector.letterbox import letterbox_for

    lib = L.lib()
    dev = torch.device("cuda:0")
    n = levels[0].shape[0]
    keep_alive = []
    lv = (L.WtHeadLevel * len(levels))()
    total = 0
    for i, t in enumerate(levels):
        h, w = t.shape[2:]
        box = t[:, :64].permute(0, 2, 3, 1).contiguous().to(dev)          # [n][hw][64] f32
        logit = t[:, 64].reshape(n, -1).contiguous().to(dev)
        keep_alive += [box, logit]
        lv[i] = L.WtHeadLevel(box.data_ptr(), None, logit.data_ptr(), h, w, (8, 16, 32)[i], L.WT_DT_F32, 0, None, 0.0)
        total += h * w
    gain, pad_x, pad_y = Y.scale_params(net_hw, img_hw)
    pp = L.WtPostParams(conf, iou, max_det, net_hw[1], net_hw[0], img_hw[1], img_hw[0], gain, pad_x, pad_y)
    out = torch.zeros((n, max_det, 6), dtype=torch.float32, device=dev)
    cnt = torch.zeros((n,), dtype=torch.int32, device=dev)
    scratch = torch.zeros(lib.wt_post_scratch_bytes(n, total), dtype=torch.uint8, device=dev)
    L.check(lib.wt_decode_nms(lv, len(levels), n, C.byref(pp), out.data_ptr(), cnt.data_ptr(), scratch.data_ptr(), 0),
            "wt_decode_nms")
    torch.cuda.synchronize()
    return out.cpu().numpy(), cnt.cpu().numpy()


def oracle_post(levels, net_hw, img_hw, conf, iou, max_det):
    pred = Y.decode_head(levels)
    res = Y.non_max_suppression(pred, conf, iou, max_det)
    out = []
    for rows, idx in res:
        rows = rows.clone()
        if rows.shape[0]:
            rows[:, :4] = Y.scale_boxes(net_hw, rows[:, :4], img_hw)
        out.append((rows.numpy(), idx.numpy()))
    return out


@pytest.mark.parametrize("net,img,max_det,conf", [
    ((384, 384), (360, 360), 1, 0.1), ((384, 384), (360, 360), 300, 0.1), ((640, 640), (640, 640), 1, 0.1),
    ((640, 640), (640, 640), 300, 0.1), ((640, 640), (360, 360), 300, 0.25), ((384, 640), (1080, 1920), 300, 0.1),
])
def test_kept_indices_bit_exact(net, img, max_det, conf):
    hw = [(net[0] // s, net[1] // s) for s in (8, 16, 32)]
    levels = random_head(5, hw, seed=net[0] + max_det)
    got, cnt = run_post(levels, net, img, conf, 0.7, max_det)
    want = oracle_post(levels, net, img, conf, 0.7, max_det)
    for i, (rows, idx) in enumerate(want):
        assert cnt[i] == len(idx), f"image {i}: kept {cnt[i]} vs oracle {len(idx)}"
        assert np.array_equal(got[i, :cnt[i], 5].astype(np.int64), idx), f"image {i}: kept anchors differ"
        if len(idx):
            assert np.abs(got[i, :cnt[i], :4] - rows[:, :4]).max() < 2e-3
            assert np.abs(got[i, :cnt[i], 4] - rows[:, 4]).max() < 1e-6


def test_no_candidates_and_all_candidates():
    hw = [(48, 48), (24, 24), (12, 12)]
    low = random_head(2, hw, seed=1, logit_mean=-9.0, logit_std=0.1)
    got, cnt = run_post(low, (384, 384), (360, 360), 0.1, 0.7, 1)
    assert cnt.tolist() == [0, 0]                                   # the reference emits NaN rows here
    high = random_head(2, hw, seed=2, logit_mean=1.0, logit_std=1.0)   # every one of the 3024 anchors passes
    got, cnt = run_post(high, (384, 384), (360, 360), 0.05, 0.7, 300)
    want = oracle_post(high, (384, 384), (360, 360), 0.05, 0.7, 300)
    for i, (rows, idx) in enumerate(want):
        assert cnt[i] == len(idx) and np.array_equal(got[i, :cnt[i], 5].astype(np.int64), idx)


def test_confidence_ties_resolve_to_lowest_anchor():
    hw = [(48, 48), (24, 24), (12, 12)]
    levels = random_head(1, hw, seed=3, logit_mean=-6.0, logit_std=0.01)
    for lv, pos in ((0, (5, 7)), (0, (30, 2)), (1, (3, 3))):
        levels[lv][0, 64, pos[0], pos[1]] = 0.5          # three exactly equal top scores
        levels[lv][0, :64, pos[0], pos[1]] = 0.0
    got, cnt = run_post(levels, (384, 384), (360, 360), 0.1, 0.7, 1)
    want = oracle_post(levels, (384, 384), (360, 360), 0.1, 0.7, 1)
    assert cnt[0] == 1 and int(got[0, 0, 5]) == int(want[0][1][0]) == 5 * 48 + 7


@pytest.mark.parametrize("max_det", [1, 300])
def test_box_conv_on_survivors_equals_box_logits(max_det):
    """wt_head_level.box_feat: the decode kernel evaluates the box branch's last 1x1 conv only for anchors that
    pass the confidence filter.  Fed with bf16 features + bf16 weights it must keep the same anchors, in the same
    order, as the plain path fed with the logits those features give (fp32 product of the same bf16 values)."""
    from wtracker_b200 import _lib as L

    lib = L.lib()
    dev = torch.device("cuda:0")
    net, img, n = (384, 384), (360, 360), 3
    hw = [(net[0] // s, net[1] // s) for s in (8, 16, 32)]
    g = torch.Generator().manual_seed(11)
    keep, lv_feat, levels = [], (L.WtHeadLevel * 3)(), []
    total = 0
    for i, (h, w) in enumerate(hw):
        feat = (torch.randn(n, h * w, 64, generator=g)).to(torch.bfloat16)
        wgt = (torch.randn(64, 64, generator=g) * 0.35).to(torch.bfloat16)
        bias = torch.randn(64, generator=g)
        logit = torch.randn(n, h * w, generator=g) * 1.2 - 3.0
        box = feat.float() @ wgt.float().T + bias                              # [n, hw, 64] f32
        levels.append(torch.cat([box.permute(0, 2, 1).reshape(n, 64, h, w), logit.reshape(n, 1, h, w)], 1))
        d = [feat.contiguous().to(dev), wgt.contiguous().to(dev), bias.to(dev), logit.contiguous().to(dev)]
        keep += d
        lv_feat[i] = L.WtHeadLevel(None, None, d[3].data_ptr(), h, w, (8, 16, 32)[i], L.WT_DT_F32, 0, None, 0.0,
                                   d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), 64)
        total += h * w
    gain, pad_x, pad_y = Y.scale_params(net, img)
    pp = L.WtPostParams(0.1, 0.7, max_det, net[1], net[0], img[1], img[0], gain, pad_x, pad_y)
    out = torch.zeros((n, max_det, 6), dtype=torch.float32, device=dev)
    cnt = torch.zeros((n,), dtype=torch.int32, device=dev)
    scratch = torch.zeros(lib.wt_post_scratch_bytes(n, total), dtype=torch.uint8, device=dev)
    L.check(lib.wt_decode_nms(lv_feat, 3, n, C.byref(pp), out.data_ptr(), cnt.data_ptr(), scratch.data_ptr(), 0),
            "wt_decode_nms")
    torch.cuda.synchronize()
    got, gcnt = out.cpu().numpy(), cnt.cpu().numpy()
    ref, rcnt = run_post(levels, net, img, 0.1, 0.7, max_det)
    assert np.array_equal(gcnt, rcnt) and gcnt.sum() > 0
    for i in range(n):
        assert np.array_equal(got[i, :gcnt[i], 5], ref[i, :gcnt[i], 5])
        assert np.abs(got[i, :gcnt[i], :4] - ref[i, :gcnt[i], :4]).max() < 1e-2
        assert np.array_equal(got[i, :gcnt[i], 4], ref[i, :gcnt[i], 4])


def test_box_and_box_feat_are_exclusive():
    from wtracker_b200 import _lib as L

    lv = (L.WtHeadLevel * 1)(L.WtHeadLevel(None, None, None, 4, 4, 8, L.WT_DT_F32, 0, None, 0.0))
    pp = L.WtPostParams(0.1, 0.7, 1, 32, 32, 32, 32, 1.0, 0.0, 0.0)
    t = torch.zeros(1024, dtype=torch.uint8, device="cuda:0")
    assert L.lib().wt_decode_nms(lv, 1, 1, C.byref(pp), t.data_ptr(), t.data_ptr(), t.data_ptr(), 0) != 0
    assert b"box" in L.lib().wt_last_error()
