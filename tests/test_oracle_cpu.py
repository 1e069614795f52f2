"""The oracle restatements against (a) the golden vectors made by the real reference and (b) the
real third-party building blocks that are installed here (cv2.resize, torchvision.ops.nms)."""
import numpy as np
import pytest
import torch

from oracle import metrics_ref, preprocess_ref, resmlp_ref
from oracle import yolov8_ref as Y
from wtracker_b200.paths import RESMLP_100 as MODEL_100, RESMLP_200 as MODEL_200
from wtracker_b200.neural.mlp import load_worm_predictor


@pytest.mark.parametrize("tag,path", [("100", MODEL_100), ("200", MODEL_200)])
def test_resmlp_oracle_matches_reference_kat(golden, tag, path):
    model = load_worm_predictor(path)
    y = resmlp_ref.resmlp_forward(model, golden[f"resmlp_{tag}_x"])
    ref = golden[f"resmlp_{tag}_y"]
    assert np.allclose(y, ref, rtol=1e-4, atol=1e-4), np.abs(y - ref).max()


def test_metric_oracle_bit_exact(golden):
    e = metrics_ref.bbox_error(golden["metric_worm"], golden["metric_mic"])
    assert np.array_equal(e, golden["metric_bbox_error"], equal_nan=True)
    m = metrics_ref.mse_error(golden["metric_worm"], golden["metric_mic"])
    assert np.array_equal(m, golden["metric_mse_error"], equal_nan=True)


def test_crop_oracle_matches_reference_views(golden):
    frames = golden["view_frames"]
    for i, pos in enumerate(golden["view_positions"]):
        f = frames[i % 3]
        p = (np.clip(pos[0], 0, f.shape[1] - 1), np.clip(pos[1], 0, f.shape[0] - 1))
        assert np.array_equal(preprocess_ref.crop_replicate(f, p, (36, 36)), golden["view_cam"][i])
        assert np.array_equal(preprocess_ref.crop_replicate(f, p, (5, 5)), golden["view_mic"][i])
        assert np.array_equal(preprocess_ref.crop_reference_style(f, p, (36, 36)), golden["view_cam"][i])


@pytest.mark.parametrize("shape", [(360, 360, 384, 384), (360, 360, 640, 640), (1080, 1920, 640, 360),
                                   (251, 251, 384, 384), (640, 640, 320, 320), (100, 37, 64, 173), (29, 29, 31, 31)])
def test_resize_oracle_bit_exact_vs_cv2(shape):
    import cv2

    sh, sw, nw, nh = shape
    img = np.random.default_rng(sum(shape)).integers(0, 256, (sh, sw), dtype=np.uint8)
    ref = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(preprocess_ref.resize_linear_u8(img, (nw, nh)), ref)


def test_letterbox_oracle_vs_cv2_pipeline():
    rng = np.random.default_rng(3)
    for hw, imgsz in (((360, 360), 384), ((640, 640), 640), ((1080, 1920), 640), ((251, 251), 384), ((300, 500), 384)):
        img = rng.integers(0, 256, hw, dtype=np.uint8)
        g = Y.letterbox_geometry(hw, imgsz)
        assert g.dst_h % 32 == 0 and g.dst_w % 32 == 0
        assert np.array_equal(preprocess_ref.letterbox_u8(img, g), Y.letterbox_cv2(img, imgsz))


def test_nms_restatement_matches_torchvision():
    import torchvision

    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 200, 1500):
        xy = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 120 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        scores[n // 3:n // 3 + 3] = scores[0]   # ties
        keep_tv = torchvision.ops.nms(boxes, scores, 0.7)
        keep_me = Y.nms_greedy(boxes, scores, 0.7)
        assert torch.equal(keep_tv, keep_me)


def test_oracle_model_shapes_and_param_count():
    from wtracker_b200.detector.weights import synthetic_state_dict

    model = Y.build_model(synthetic_state_dict(0))
    assert sum(p.numel() for p in model.parameters()) == 11_125_955 - 0  # fused: BN folded into conv bias
    with torch.no_grad():
        pred = model(torch.rand(1, 3, 64, 96))
    assert pred.shape == (1, 5, 8 * 12 + 4 * 6 + 2 * 3)
