"""Screen-band sharding across the GPUs of one box.

The draw path shards in screen space exactly like the reference shards across its 16 clusters
(/root/reference/src/Device/QuadRasterizer.cpp:46-49): pixels are independent given in-order triangles, so each rank
renders the full triangle list with renderArea = its band of rows and the finished bands are all-gathered.  Bands start on
even rows (quads are two rows tall, so LOD derivatives never straddle a seam) and have equal sizes (all_gather).
Backend-agnostic: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations


def band_rows(height: int, world: int, rank: int) -> tuple:
    """[y0, y1) of rank's band.  The height must split into `world` bands of an even number of rows."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if height % (2 * world):
        raise ValueError(f"framebuffer height {height} does not split into {world} bands of even height")
    rows = height // world
    return rank * rows, (rank + 1) * rows


def render_area(width: int, height: int, world: int, rank: int) -> tuple:
    y0, y1 = band_rows(height, world, rank)
    return (0, y0, width, y1 - y0)


def gather_bands(full, height: int, pitch_bytes: int, world: int, rank: int):
    """In-place all-gather of the finished bands: `full` is a flat uint8 torch tensor over the whole 1x image whose rows
    [y0, y1) this rank has rendered.  After the call every rank (in particular rank 0, which presents) holds the frame."""
    import torch.distributed as dist
    if world == 1:
        return full
    y0, y1 = band_rows(height, world, rank)
    mine = full[y0 * pitch_bytes: y1 * pitch_bytes]
    dist.all_gather_into_tensor(full[: height * pitch_bytes], mine)
    return full
