import torch, traceback
from wtracker_b200.detector.weights import synthetic_state_dict
from wtracker_b200.detector.engine import DetectorEngine
sd = synthetic_state_dict(0)
try:
    eng = DetectorEngine(sd, (640, 640), 640, batch=2, max_det=1)
    eng.input_view.random_(0, 255)
    torch.cuda.synchronize()
    eng.forward(2, 0, 1)
    torch.cuda.synchronize()
    print("op0 ok")
except Exception as e:
    print("ERR", repr(e)[:300])
