"""Structure of the compiled detector program (host logic, no GPU): for several network sizes every channel range an op
reads was written by an EARLIER op of the program (the engine derives its cross-lane waits from program order), the fused
forms appear where they should, and weight offsets satisfy the alignment the C ABI asks for."""
import pytest

from wtracker_b200 import _lib as L
from wtracker_b200.detector.program import build_program
from wtracker_b200.detector.weights import infer_arch, synthetic_state_dict


def reads_writes(o, p):
    rd, wr = [], []
    if o["kind"] == L.WT_OP_CONV:
        rd.append((o["src"], o["src_coff"], o["src_coff"] + o["cin"]))
        if o["res"] >= 0:
            rd.append((o["res"], o["res_coff"], o["res_coff"] + o["cout"]))
        if o.get("add_buf", -1) >= 0:
            rd.append((o["add_buf"], o["add_coff"], o["add_coff"] + o["cout"]))
        cat = o.get("chain_w_off", -1) >= 0 and o.get("cat_buf", -1) >= 0
        if cat:
            rd.append((o["cat_buf"], o["cat_coff"], o["cat_coff"] + o["cat_c"]))
        if o.get("dot_off", -1) >= 0:
            wr.append((o["dst"], 0, 1))
        else:
            wr.append((o["dst"], o["dst_coff"], o["dst_coff"] + (o["chain_cout"] if cat else o["cout"])))
    elif o["kind"] == L.WT_OP_CONV0:
        rd.append((o["src"], 0, 1))
        wr.append((o["dst"], o["dst_coff"], o["dst_coff"] + o["cout"]))
    elif o["kind"] == L.WT_OP_SPPF_POOL:
        rd.append((o["src"], o["src_coff"], o["src_coff"] + o["cin"]))
        wr.append((o["dst"], o["dst_coff"], o["dst_coff"] + 3 * o["cin"]))
    return rd, wr


@pytest.mark.parametrize("hw", [(640, 640), (384, 384), (384, 640), (352, 352)])
@pytest.mark.parametrize("chain", [True, False])
def test_program_is_topologically_ordered(hw, chain):
    sd = synthetic_state_dict(0)
    p = build_program(sd, infer_arch(sd), hw[0], hw[1], chain=chain)
    written = {0: [(0, 1)]}                                  # buffer 0 = the u8 network input
    for i, o in enumerate(p.ops):
        rd, wr = reads_writes(o, p)
        for buf, lo, hi in rd:
            covered = sorted(written.get(buf, []))
            c = lo
            for a, b in covered:                             # the read range must be a union of earlier writes
                if a <= c < b:
                    c = b
            assert c >= hi, f"op {i} {o['name']} reads buffer {p.buf_names[buf]}[{lo}:{hi}] before it is written"
        for buf, lo, hi in wr:
            assert hi <= p.bufs[buf][2], (o["name"], p.buf_names[buf])
            written.setdefault(buf, []).append((lo, hi))
        assert o.get("lane", 0) in (0, 1)
        if o["kind"] == L.WT_OP_CONV:
            assert o["w_off"] % 16 == 0 and o["b_off"] % 4 == 0
            if o.get("chain_w_off", -1) >= 0:
                assert o["chain_w_off"] % 16 == 0 and o["chain_b_off"] % 4 == 0
    names = [o["name"] for o in p.ops]
    if chain:
        assert "model.1>model.2.cv1" in names and "model.3>model.4.cv1" in names
        assert ("model.2.m.0.cv2>model.2.cv2" in names) == (hw[0] % 64 == 0 and hw[1] % 32 == 0)
        assert "m1" not in p.taps and "m3" not in p.taps
    else:
        assert not any(">" in n for n in names) and len(p.ops) == 57
    assert sum(o.get("lane", 0) for o in p.ops) == 6         # the heads of levels 0 and 1 run on the side stream
    assert len(p.head) == 3 and p.total_anchors == sum((hw[0] // s) * (hw[1] // s) for s in (8, 16, 32))
