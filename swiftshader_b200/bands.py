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


class Group:
    """The GPUs of one box render ONE frame (swcu_group_*): rank r sets up triangles [r n / world, (r + 1) n / world) of every
    binned draw for the whole frame and stores each record, its bin counts and its big-list entry into the work buffers of the
    rank(s) whose band the triangle touches (CUDA IPC mappings, NVLink); one flag barrier per draw, then every rank bins and
    rasterises its band.  torch.distributed only carries the 64-byte handles at start-up.

    Every rank must then issue the same binned draws in the same order with render_area = its band (``render_area`` above)."""

    def __init__(self, dev, scene_or_dims, world: int, rank: int, max_primitives: int, max_slots: int = 6, max_samples: int = 4):
        import ctypes as C
        import torch.distributed as dist
        from . import capi
        self.dev, self.world, self.rank = dev, world, rank
        w, h = scene_or_dims if isinstance(scene_or_dims, tuple) else (scene_or_dims.width, scene_or_dims.height)
        g = capi.GroupDesc(C.sizeof(capi.GroupDesc), rank, world, max_primitives, max_slots, max_samples, w, h)
        handle = (C.c_ubyte * 64)()
        err = ""
        try:
            dev.check(dev.lib.swcu_group_reserve(dev.ctx, C.byref(g), handle))
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        got = [None] * world
        dist.all_gather_object(got, (bytes(handle), err))
        errs = [e for (_, e) in got if e]
        if errs:
            dev.lib.swcu_group_detach(dev.ctx)
            raise RuntimeError("group reserve failed: " + "; ".join(errs))
        blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(h_ for (h_, _) in got))
        err = ""
        try:
            dev.check(dev.lib.swcu_group_attach(dev.ctx, blob))
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        errs = [None] * world
        dist.all_gather_object(errs, err)
        if any(errs):
            dev.lib.swcu_group_detach(dev.ctx)
            raise RuntimeError("group attach failed: " + "; ".join(e for e in errs if e))
        dist.barrier()

    def close(self):
        import torch.distributed as dist
        self.dev.sync()
        dist.barrier()  # nobody unmaps while a peer may still be writing
        self.dev.check(self.dev.lib.swcu_group_detach(self.dev.ctx))
        dist.barrier()


class PeerGather:
    """Finished bands reach rank 0's frame by STORES OVER NVLINK, not by a collective: rank 0 exports the CUDA IPC handles of
    its presentable 1x frames (one or a small ring of them) and of a flag array, every other rank maps them and lets its end-of-pass
    resolve (4x) or band copy (1x) write straight into the frame of the current slot; counters that only grow order the frames
    (flags[16 s + r] = "band r of the frame in slot s has arrived", written by rank r; flags[16 s] = "the frame in slot s has been
    consumed", written by rank 0 and polled by the others over NVLink before they overwrite it).  With a ring of frames rank 0 can
    still be downloading frame i while the bands of frame i + 1 land in the next slot.  torch.distributed only carries the handles.

    `dev` is a swiftshader_b200.scene.Device, `final_host` the numpy array (or list of arrays: the ring) that keys the frame(s)."""

    def __init__(self, dev, final_host, height: int, pitch_bytes: int, world: int, rank: int):
        import ctypes as C
        import numpy as np
        import torch.distributed as dist
        self.dev, self.world, self.rank = dev, world, rank
        self.height, self.pitch = height, pitch_bytes
        frames = list(final_host) if isinstance(final_host, (list, tuple)) else [final_host]
        self.slots = len(frames)
        assert 1 <= self.slots <= 4 and world <= 16
        self.frame_no = 0
        self.peer_frames, self.peer_flags = [], None
        self._opened = []
        self.flags = np.zeros(64, dtype=np.uint32)  # rank 0's copy is the live one
        dev.register(self.flags)
        dev.upload(self.flags)
        dev.sync()
        # Every step that can fail (CUDA IPC may be unavailable) is agreed on by all ranks before anyone waits for anyone:
        # a failure raises on every rank instead of leaving the others in a collective.
        payload = [None]
        if rank == 0:
            try:
                hs = []
                for arr in frames + [self.flags]:
                    h = (C.c_ubyte * 64)()
                    off = C.c_uint64(0)
                    dev.check(dev.lib.swcu_ipc_export(dev.ctx, arr.ctypes.data, h, C.byref(off)))
                    hs.append((bytes(h), int(off.value)))
                payload = [hs]
            except Exception as e:  # noqa: BLE001
                payload = [f"export failed: {e}"]
        dist.broadcast_object_list(payload, src=0)
        if isinstance(payload[0], str):
            dev.unregister(self.flags)
            raise RuntimeError(payload[0])
        err = ""
        if rank != 0:
            try:
                hs = payload[0]
                for (hf, of) in hs[:-1]:
                    pf = self._open(hf) + of
                    # adopt the mappings so that addresses inside them are accepted as attachments / flags
                    dev.check(dev.lib.swcu_mem_register_device(dev.ctx, pf, height * pitch_bytes))
                    self.peer_frames.append(pf)
                (hg, og) = hs[-1]
                self.peer_flags = self._open(hg) + og
                dev.check(dev.lib.swcu_mem_register_device(dev.ctx, self.peer_flags, 64 * 4))
            except Exception as e:  # noqa: BLE001
                err = f"rank {rank}: {e}"
        errs = [None] * world
        dist.all_gather_object(errs, err)
        if any(errs):
            raise RuntimeError("peer mapping failed: " + "; ".join(e for e in errs if e))
        dist.barrier()

    def _open(self, handle: bytes) -> int:
        import ctypes as C
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        base = C.c_void_p()
        self.dev.check(self.dev.lib.swcu_ipc_open(self.dev.ctx, buf, C.byref(base)))
        self._opened.append(base.value)
        return base.value

    @property
    def slot(self) -> int:
        """Slot of the frame begun last."""
        return (self.frame_no - 1) % self.slots

    def _round(self) -> int:
        """How many times the current slot has been used, this frame included (the value its flags are waiting for)."""
        return (self.frame_no - 1) // self.slots + 1

    def band_destination(self, fmt: int, width: int, slot: int = None):
        """Attachment describing my band inside rank 0's frame of `slot` (ranks > 0; default: the current slot)."""
        from . import capi
        y0, y1 = band_rows(self.height, self.world, self.rank)
        base = self.peer_frames[self.slot if slot is None else slot]
        return capi.Attachment(base + y0 * self.pitch, fmt, self.pitch, 0, width, y1 - y0, 0)

    def begin_frame(self):
        """Before this rank overwrites its band of rank 0's frame: the frame that was in this slot must have been consumed."""
        self.frame_no += 1
        if self.rank != 0 and self._round() > 1:
            self.dev.check(self.dev.lib.swcu_wait_flags(self.dev.ctx, self.peer_flags + 64 * self.slot, 0, 1, self._round() - 1))

    def band_done(self):
        """After the band's resolve / copy has been enqueued: announce it (ranks > 0), or wait for all bands (rank 0)."""
        if self.rank != 0:
            self.dev.check(self.dev.lib.swcu_signal(self.dev.ctx, self.peer_flags + 64 * self.slot + 4 * self.rank, self._round()))
        else:
            self.dev.check(self.dev.lib.swcu_wait_flags(self.dev.ctx, self.flags.ctypes.data + 64 * self.slot, 1, self.world - 1, self._round()))

    def frame_consumed(self):
        """Rank 0, once whoever presents the frame is done with it (enqueued on the same stream)."""
        if self.rank == 0:
            self.dev.check(self.dev.lib.swcu_signal(self.dev.ctx, self.flags.ctypes.data + 64 * self.slot, self._round()))

    def close(self):
        import torch.distributed as dist
        self.dev.sync()
        dist.barrier()
        if self.rank != 0:
            for pf in self.peer_frames:
                self.dev.check(self.dev.lib.swcu_mem_unregister(self.dev.ctx, pf))
            self.dev.check(self.dev.lib.swcu_mem_unregister(self.dev.ctx, self.peer_flags))
            for b in self._opened:
                self.dev.lib.swcu_ipc_close(self.dev.ctx, b)
        self.dev.unregister(self.flags)
        dist.barrier()
