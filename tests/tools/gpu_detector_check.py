"""Diagnostics: CUDA detector vs the fp32 torch oracle on synthetic camera views (run on a GPU box)."""
import sys, time
import numpy as np, torch
from wtracker_b200 import synth
from wtracker_b200.detector.weights import synthetic_state_dict
from wtracker_b200.detector.engine import DetectorEngine
from oracle import yolov8_ref as O
from oracle.preprocess_ref import letterbox_u8

view, imgsz, n = int(sys.argv[1]) if len(sys.argv) > 1 else 360, int(sys.argv[2]) if len(sys.argv) > 2 else 384, 4
impl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sd = synthetic_state_dict(0)
track = synth.worm_track(2000, 0)
views = []
for i in range(n):
    f = synth.render_frame(i * 400, track, 0)
    pos = (int(track[i * 400, 0]) + 20 * i, int(track[i * 400, 1]) - 10 * i)
    views.append(np.ascontiguousarray(synth.camera_view(f, pos, view)))
model = O.build_model(sd)
eng = DetectorEngine(sd, (view, view), imgsz, batch=n, max_det=1, conv_impl=impl)
t = time.time(); boxes, counts = eng.detect_views(views); torch.cuda.synchronize(); print("gpu detect", time.time() - t)
# pre
lbimg = np.stack([letterbox_u8(v, eng.lb) for v in views])
gin = eng.input_view.cpu().numpy()
print("pre exact:", np.array_equal(lbimg, gin), "maxdiff", np.abs(lbimg.astype(int) - gin.astype(int)).max())
x = O.preprocess(views, imgsz)
taps = {}
with torch.no_grad():
    feats = model.features(x, taps)
for name, (bid, coff, c) in eng.program.taps.items():
    g = eng.buffer_tensor(bid, n)[..., coff:coff + c].float().permute(0, 3, 1, 2).cpu()
    r = taps[name]
    err = (g - r).abs()
    print(f"{name:4s} ref std {r.std():.4f} max {r.abs().max():.3f} | abs err max {err.max():.4f} mean {err.mean():.5f} rel(mean/std) {err.mean()/r.std():.4f}")
for lvl, h in enumerate(eng.program.head):
    cb = model.model[22].cv2[lvl][2]
    gb = (eng.buffer_tensor(h["box_feat"], n).float().cpu() @ cb.weight.view(64, -1).to(torch.bfloat16).float().T
          + cb.bias.detach()).permute(0, 3, 1, 2)
    rb = feats[lvl][:, :64]
    e = (gb - rb).abs()
    print(f"head{lvl} box logits ref std {rb.std():.3f} | err max {e.max():.4f} mean {e.mean():.5f}")
    glog = eng.buffer_tensor(h["cls_logit"], n)[..., 0].cpu()
    rl = feats[lvl][:, 64]
    e = (glog - rl).abs()
    print(f"head{lvl} cls logit ref mean {rl.mean():.3f} std {rl.std():.3f} max {rl.max():.3f} | err max {e.max():.4f} mean {e.mean():.5f}")
orc = O.YoloOracle(model, imgsz, max_det=1)
res = orc.detect(views)
for i, (rows, idx) in enumerate(res):
    print(i, "oracle", rows.numpy().round(3), idx.numpy(), "| gpu", counts[i], boxes[i, :counts[i]].round(3))
del eng
eng = DetectorEngine(sd, (view, view), imgsz, batch=n, max_det=300, conv_impl=impl)
boxes, counts = eng.detect_views(views)
res = O.YoloOracle(model, imgsz, max_det=300).detect(views)
for i, (rows, idx) in enumerate(res):
    gi = boxes[i, :counts[i], 5].astype(int)
    oi = idx.numpy()
    same = len(gi) == len(oi) and (gi == oi).all()
    common = len(set(gi) & set(oi))
    print(i, "max_det=300: oracle kept", len(oi), "gpu kept", len(gi), "identical order:", same, "common", common)
    if len(gi) == len(oi) and same:
        d = np.abs(boxes[i, :counts[i], :4] - rows[:, :4].numpy())
        print("   box abs diff max", d.max(), "conf diff max", np.abs(boxes[i, :counts[i], 4] - rows[:, 4].numpy()).max())
    else:
        print("   oracle", oi[:12], rows[:12, 4].numpy().round(4)); print("   gpu   ", gi[:12], boxes[i, :12, 4].round(4))
