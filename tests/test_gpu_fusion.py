"""Graph-level fusions must not change a bit: the detector program with chained convs, the concat chain and launch
lanes (54 ops) against the same program built without them (57 ops, one stream) on the same input — every tapped
feature map, the box features and the class logits are compared exactly.  (Per-kernel versions of the same claim:
wt_selftest_conv_chain / wt_selftest_conv_cat, tools/gpu_conv_selftest.py.)"""
import pytest
import torch

from gpu_common import synthetic_sd, views_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("view,imgsz", [(360, 384), (640, 640)])
def test_fused_program_is_bit_identical_to_the_unfused_one(view, imgsz):
    from wtracker_b200 import _lib as L
    from wtracker_b200.detector.engine import DetectorEngine

    views = views_for(view, 4)
    fused = DetectorEngine(synthetic_sd(), (view, view), imgsz, batch=4, max_det=1)
    plain = DetectorEngine(synthetic_sd(), (view, view), imgsz, batch=4, max_det=1, fuse=False)
    assert len(fused.program.ops) == 54 and len(plain.program.ops) == 57
    assert any(o.get("cat_buf", -1) >= 0 for o in fused.program.ops)
    assert sum(1 for o in fused.program.ops if o.get("chain_w_off", -1) >= 0 and o["kind"] == L.WT_OP_CONV) == 3
    assert not any(o.get("chain_w_off", -1) >= 0 and o["kind"] == L.WT_OP_CONV for o in plain.program.ops)
    b1, c1 = fused.detect_views(views)
    b2, c2 = plain.detect_views(views)
    for name, (bid, coff, c) in fused.program.taps.items():
        pb, pc, _ = plain.program.taps[name]
        a = fused.buffer_tensor(bid, 4)[..., coff:coff + c]
        b = plain.buffer_tensor(pb, 4)[..., pc:pc + c]
        assert torch.equal(a, b), name
    for hf, hp in zip(fused.program.head, plain.program.head):
        assert torch.equal(fused.buffer_tensor(hf["box_feat"], 4), plain.buffer_tensor(hp["box_feat"], 4))
        assert torch.equal(fused.buffer_tensor(hf["cls_logit"], 4), plain.buffer_tensor(hp["cls_logit"], 4))
    assert (c1 == c2).all() and (b1 == b2).all()
