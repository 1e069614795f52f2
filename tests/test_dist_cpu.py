"""Multi-GPU plan on CPU: frame-range / experiment sharding and the final gather of the per-frame
result table, exercised with the gloo backend at world_size 2 (the N>1 path of bench.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wtracker_b200.sharding import frame_range, gather_result_table, result_table_from


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = frame_range(total, rank, world)
    idx = torch.arange(lo, hi)
    boxes = torch.stack([idx * 1.0, idx * 2.0, idx * 0 + 3.0, idx * 0 + 4.0, idx * 0.001, idx * 7.0], 1)
    counts = (idx % 5 != 0).int()
    table = result_table_from(boxes.view(-1, 1, 6), counts, idx)
    full = gather_result_table(table, total)
    if rank == 0:
        q.put(full.numpy())
    dist.destroy_process_group()


def test_frame_ranges_partition_the_work():
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [frame_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    total = 101
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full.shape == (total, 8)
    idx = np.arange(total)
    assert np.array_equal(full[:, 6], idx.astype(np.float32))          # frame index column, in order
    valid = (idx % 5 != 0)
    assert np.array_equal(full[:, 7] > 0, valid)
    assert np.allclose(full[valid, 0], idx[valid]) and np.isnan(full[~valid, :4]).all()


def _worker_rows(rank, world, port, total, q):
    """The offline driver's table: int32 rows (wt_result_rows layout, float columns bit-cast) through the same gather."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = frame_range(total, rank, world)
    rows = np.zeros((hi - lo, 8), dtype=np.int32)
    f = np.arange(lo, hi)
    rows.view(np.float32)[:, 0] = f * 0.5
    rows.view(np.float32)[:, 4] = np.where(f % 7 == 0, np.nan, 0.25)
    rows[:, 5] = np.where(f % 7 == 0, -1, f * 3)
    rows[:, 6] = f
    rows[:, 7] = f % 7 != 0
    full = gather_result_table(torch.from_numpy(rows), total)
    if rank == 1:
        q.put(full.numpy())
    dist.destroy_process_group()


def test_gather_int32_rows_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    total = 77                                   # odd: the shares differ by one row
    procs = [ctx.Process(target=_worker_rows, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert full.dtype == np.int32 and full.shape == (total, 8)
    f = np.arange(total)
    assert np.array_equal(full[:, 6], f) and np.array_equal(full[:, 7], (f % 7 != 0).astype(np.int32))
    assert np.array_equal(full.view(np.float32)[:, 0], (f * 0.5).astype(np.float32))
    assert np.array_equal(np.isnan(full.view(np.float32)[:, 4]), f % 7 == 0)      # NaN bit patterns survive the gather
    assert np.array_equal(full[:, 5], np.where(f % 7 == 0, -1, f * 3))
