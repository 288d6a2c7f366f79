"""CPU, world_size 2 over gloo: the band split + all-gather that bench.py runs over NCCL.  Each rank renders ITS band of a
scene with the CPU oracle (test infrastructure; the product renders the band on its GPU), the bands are all-gathered in
place, and both ranks must hold the same frame as a single full-frame render."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import swref
        from swiftshader_b200 import bands, workloads
        sc = workloads.small(name).scene
        H, W = sc.height, sc.width
        area = bands.render_area(W, H, world, rank)
        att = swref.render_oracle(sc, render_area=area)
        img = swref.resolve_oracle(sc, att) if sc.samples > 1 else att["color"][0]
        full = torch.from_numpy(np.ascontiguousarray(img[:H]).reshape(-1).copy())
        bands.gather_bands(full, H, W * 4, world, rank)
        whole = swref.render_oracle(sc)
        ref = swref.resolve_oracle(sc, whole) if sc.samples > 1 else whole["color"][0]
        ok = np.array_equal(full.numpy().reshape(H, W, 4), ref[:H])
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c4", "c5"])
def test_two_ranks_reassemble_the_frame(name):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + (0 if name == "c4" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


def test_band_rows():
    from swiftshader_b200 import bands
    assert bands.band_rows(2160, 8, 0) == (0, 270)
    assert bands.band_rows(2160, 8, 7) == (1890, 2160)
    assert bands.render_area(3840, 2160, 2, 1) == (0, 1080, 3840, 1080)
    with pytest.raises(ValueError):
        bands.band_rows(1080, 16, 0)  # 67.5 rows
    with pytest.raises(ValueError):
        bands.band_rows(100, 2, 2)
