#!/usr/bin/env python
"""bench.py — the draw hot path on N B200s, one JSON line (see the contract in the task statement / DESIGN.md §Measurement).

A "step" is one frame of the workload: every draw of the scene through swcu_draw (setup -> binning -> region raster / shade /
blend) plus, for the multisampled workload, the end-of-pass resolve.  At N > 1 the GPUs form a group (swcu_group_*): every rank
sets up 1/N of the triangles for the whole frame and stores the records into the owning rank's buffers over NVLink, renders its
screen band, and the finished bands reach rank 0 by stores over NVLink (default) or an NCCL all-gather (SWCU_GATHER=nccl).
Attachments stay resident; the clear is outside the timed region, like the reference harness' LOAD pass (oracle/refrender.cpp --time).

The headline workload is C4 (the largest single-GPU configuration of BASELINE.json); every run additionally measures C5 (the
multi-GPU configuration: 10 M triangles at 8K) under "secondary", and checks the frame rank 0 holds against the sha256 of the
reference ICD's own render of the same scene ("frame_hash_ok").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl cuda|reference] [--no-secondary]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gpixels/s shaded+blended (also Mtris/s; HBM GB/s vs roofline)"
DEFAULT_WORKLOAD = "c4"
SECONDARY = {"c4": "c5"}  # the multi-GPU configuration of BASELINE.json rides along with the headline one


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernels_hash() -> str:
    """sha256 (16 hex digits) of the kernel sources: ties profiles/traffic.json to the code it was captured from."""
    h = hashlib.sha256()
    for f in ("kernels.cuh", "draw.cu", "swcu_internal.h"):
        h.update(open(os.path.join(ROOT, "swiftshader_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    return min(os.cpu_count() or 1, 16)  # SwiftShader: min(logical CPUs, 16) workers (src/System/SwiftConfig.cpp:128)


def reference_time(wl, frames: int, warmup: int) -> dict:
    """Times the reference's own CPU implementation (oracle/_ref: the unmodified ICD driven by oracle/refrender) on the
    box's host cores.  This is the one place bench.py executes anything under oracle/."""
    from oracle import swref
    if not swref.reference_available():
        raise RuntimeError("oracle/_ref is missing (built by __graft_entry__.build() where /root/reference exists)")
    out = swref.render_reference(wl.scene, time_frames=frames, warmup=max(1, warmup))
    return out["timing"]


EST_REF_MS = {"c1": 3, "c2": 4, "c3": 15, "c4": 700, "c5": 4000}


def run_reference(args, rank: int):
    if rank != 0:
        return
    from swiftshader_b200 import workloads
    wl = workloads.WORKLOADS[args.workload]()
    # bounded sample: the reference renders the whole frame K times; K is clamped so the run ends within minutes
    frames = max(1, min(args.steps, int(120000 / EST_REF_MS[args.workload])))
    warm = max(1, min(args.warmup, 3))
    t = reference_time(wl, frames, warm)
    ms = t["mean_ms"]
    val = wl.covered_pixels / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gpixels/s", "n_gpus": args.gpus, "steps": frames, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
        "mtris_per_s": wl.triangles / (ms * 1e-3) / 1e6,
        "config": {"workload": wl.name, "description": wl.description, "threads": host_threads(), "device": t.get("device")},
        "cpu_baseline": {"value": val, "unit": "Gpixels/s", "cores": host_threads(), "kind": "reference",
                         "sample": f"{frames} whole frames of {wl.name} (draw-only LOAD pass + resolve), after {warm} warm-up, mean"},
        "e2e": {"value": val, "unit": "Gpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def run_workload(name: str, steps: int, warmup: int, rank: int, N: int, local_rank: int, stream, with_e2e: bool = True, sustain_s: float = 1.0) -> dict:
    """Everything bench.py measures for one workload; returns the pieces of the JSON line (only rank 0's copy is printed)."""
    import ctypes as C
    import dataclasses
    import numpy as np
    import torch
    import torch.distributed as dist
    from swiftshader_b200 import bands, capi, workloads
    from swiftshader_b200.scene import Device, Frame

    wl = workloads.WORKLOADS[name]()
    sc = wl.scene
    H, W = sc.height, sc.width
    band = bands.band_rows(H, N, rank)
    area = bands.render_area(W, H, N, rank)
    pitch = W * 4
    H2 = sc.padded_height()
    dev_s = f"cuda:{local_rank}"

    dev = Device(local_rank)
    dev.set_stream(stream.cuda_stream)
    # N > 1: the group shards the setup of every draw by triangle range (SWCU_GROUP=0: every rank sets all triangles up itself)
    group = None
    if N > 1 and os.environ.get("SWCU_GROUP", "1") != "0":
        group = bands.Group(dev, sc, N, rank, max(d.primitive_count() for d in sc.draws), 6, sc.samples)
    frame = Frame(dev, sc, render_area=area)
    frame.upload_inputs()
    frame.clear()

    # resolve only my band: attachments re-based to the band's first row
    def band_att(host_arr, slice_b):
        return capi.Attachment(host_arr.ctypes.data + band[0] * pitch, sc.colorFormat, pitch, slice_b, W, band[1] - band[0], 0)

    src_b = band_att(frame.att["color"], H2 * pitch)
    # The presentable 1x frames: a ring of three (MSAA: resolve targets; 1x at N > 1: copy targets on rank 0), so that frame i can
    # still be on its way to the host while frame i + 1 is rendered into the next slot.  1x at N = 1: the colour attachment itself.
    ring_n = 3 if (sc.samples > 1 or N > 1) else 1
    finals = []
    if sc.samples > 1:
        finals.append(frame.resolved)
    elif N > 1 and rank == 0:
        finals.append(np.zeros((1, H2, W, 4), dtype=np.uint8))
        dev.register(finals[0], upload=False)
    while finals and len(finals) < ring_n and (rank == 0 or N == 1):
        extra = np.zeros_like(finals[0])
        dev.register(extra, upload=False)
        finals.append(extra)
    final_imgs = [f[0] for f in finals] if finals else [frame.att["color"][0]]
    final_atts = [band_att(f, H2 * pitch) for f in finals]

    # N > 1: how the finished bands reach rank 0.  "peer" (default): the end-of-pass resolve (4x) / a band copy (1x) of every
    # rank stores straight into rank 0's frame of the current ring slot over NVLink (CUDA IPC mapping), ordered by flags — no
    # collective on the data path.  "nccl": in-place NCCL all-gather of the bands (SWCU_GATHER=nccl).
    gather = os.environ.get("SWCU_GATHER", "peer") if N > 1 else "none"
    pg, full = None, None
    if gather == "peer":
        try:
            pg = bands.PeerGather(dev, final_imgs if rank == 0 else [None] * ring_n, H, pitch, N, rank)
            ok = 1
        except Exception as e:  # noqa: BLE001  (e.g. CUDA IPC not permitted in this container)
            print(f"[bench] rank {rank}: peer delivery unavailable ({e}); falling back to the NCCL all-gather", file=sys.stderr)
            pg, ok = None, 0
        t_ok = torch.tensor([ok], device=dev_s)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            pg, gather = None, "nccl"
    if gather == "nccl":
        if sc.samples == 1 and N > 1:  # every rank needs a whole-frame buffer to gather into
            if not finals:
                finals.append(np.zeros((1, H2, W, 4), dtype=np.uint8))
                dev.register(finals[0], upload=False)
                final_imgs, final_atts = [finals[0][0]], [band_att(finals[0], H2 * pitch)]
        full = torch.as_tensor(_DevArr(dev.device_ptr(final_imgs[0]), H * pitch), device=dev_s)
        dev.set_option("copy_streams", 0)  # the all-gather writes the frame on the caller's stream, unseen by the library

    # Peer delivery runs on the library's hand-over stream (SWCU_HANDOVER=0: on the main stream, between the frames): a rank resolves /
    # copies its band into a LOCAL 1x band buffer (two of them, used in turn) and goes on with the next frame; the wait for the ring
    # slot, the copy of the band into rank 0's frame over NVLink and the signal follow on the second stream.  Rank 0 likewise waits
    # for the other ranks' bands (and downloads / releases the frame) beside its next frame.
    handover = pg is not None and os.environ.get("SWCU_HANDOVER", "1") != "0"
    local_atts = []
    if handover and rank != 0:
        for _ in range(2):
            lb = np.zeros((band[1] - band[0], W, 4), dtype=np.uint8)
            dev.register(lb, upload=False)
            finals.append(lb)  # (kept alive)
            local_atts.append(capi.Attachment(lb.ctypes.data, sc.colorFormat, pitch, 0, W, band[1] - band[0], 0))

    def side_join():
        """The main stream catches up with the hand-over stream (before a time stamp that is to cover whole frames)."""
        if handover:
            for k_ in range(2):
                dev.check(dev.lib.swcu_side_wait(dev.ctx, k_))

    state = {"i": 0}

    def step(present=None, descs=None):
        """One frame: draws, end-of-pass resolve / band copy into the presentable frame of the current slot, delivery to rank 0."""
        i = state["i"]
        state["i"] += 1
        for d_ in (descs if descs is not None else frame.descs):
            dev.draw(d_)
        k = i % len(final_imgs)
        if pg is not None and handover:
            k2 = i & 1
            if rank != 0:
                dev.check(dev.lib.swcu_side_wait(dev.ctx, k2))  # the delivery of frame i - 2 has read this band buffer
                if sc.samples > 1:
                    dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src_b), sc.samples, C.byref(local_atts[k2])))
                else:
                    dev.check(dev.lib.swcu_copy_image(dev.ctx, C.byref(src_b), C.byref(local_atts[k2])))
                dev.check(dev.lib.swcu_side_begin(dev.ctx))
                pg.begin_frame()  # the frame that was in this ring slot must have been consumed
                dev.check(dev.lib.swcu_copy_image(dev.ctx, C.byref(local_atts[k2]), C.byref(pg.band_destination(sc.colorFormat, W))))
                pg.band_done()    # announce the band
                dev.check(dev.lib.swcu_side_end(dev.ctx, k2))
            else:
                pg.begin_frame()
                dst = final_atts[pg.slot]
                if sc.samples > 1:
                    dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src_b), sc.samples, C.byref(dst)))
                else:
                    dev.check(dev.lib.swcu_copy_image(dev.ctx, C.byref(src_b), C.byref(dst)))
                dev.check(dev.lib.swcu_side_begin(dev.ctx))
                pg.band_done()    # the hand-over stream waits for the other ranks' bands
                if present is not None:
                    present(final_imgs[pg.slot])
                pg.frame_consumed()
                dev.check(dev.lib.swcu_side_end(dev.ctx, k2))
            return
        if pg is not None:
            pg.begin_frame()  # (ranks > 0: the frame that was in this slot must have been consumed)
            dst = pg.band_destination(sc.colorFormat, W) if rank != 0 else final_atts[pg.slot]
            if sc.samples > 1:
                dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src_b), sc.samples, C.byref(dst)))
            else:
                dev.check(dev.lib.swcu_copy_image(dev.ctx, C.byref(src_b), C.byref(dst)))
            pg.band_done()  # ranks > 0: announce; rank 0: the stream waits for the other ranks' bands
            if rank == 0:
                if present is not None:
                    present(final_imgs[pg.slot])
                pg.frame_consumed()
            return
        if sc.samples > 1:
            dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src_b), sc.samples, C.byref(final_atts[0 if gather == "nccl" else k])))
        elif gather == "nccl":
            dev.check(dev.lib.swcu_copy_image(dev.ctx, C.byref(src_b), C.byref(final_atts[0])))
        if gather == "nccl":
            bands.gather_bands(full, H, pitch, N, rank)  # NCCL all-gather, in place: my band is already at its slot
        if present is not None and rank == 0:
            present(final_imgs[0 if gather == "nccl" else k])

    def barrier():
        if N > 1:
            dist.barrier()

    out = {"wl": wl, "band_rows": band[1] - band[0], "gather": gather, "group": group is not None}
    with torch.cuda.stream(stream):
        # ---- correctness first: ONE frame from cleared attachments, assembled on rank 0, against the reference ICD's own render ----
        got = {}
        step(present=lambda img: (dev.download(img), got.setdefault("img", img)))
        dev.sync()
        torch.cuda.synchronize()
        barrier()
        if rank == 0:
            try:
                want = json.load(open(os.path.join(ROOT, "tests", "golden", "workload_hashes.json")))[name]["hashes"]["color"]
                out["frame_hash_ok"] = hashlib.sha256(np.ascontiguousarray(got["img"][:H]).tobytes()).hexdigest() == want
            except (OSError, KeyError):
                out["frame_hash_ok"] = None

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        dev.reset_stats()
        # one event per frame boundary: the contract's value is total / K; the per-frame spread (p10 / p50 / p90) is reported too
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        torch.cuda.synchronize()
        marks[0].record(stream)
        t_host0 = time.perf_counter()
        for i in range(steps):
            step()
            if i == steps - 1:
                side_join()  # the last time stamp covers the delivery of every frame
            marks[i + 1].record(stream)
        # what the host needs to ENQUEUE a frame (the calls are asynchronous): a frame cannot be faster than this
        out["host_enqueue_ms_step"] = (time.perf_counter() - t_host0) * 1e3 / steps
        torch.cuda.synchronize()
        barrier()
        ms_total = marks[0].elapsed_time(marks[-1])
        per_frame = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
        pct = lambda q: per_frame[min(len(per_frame) - 1, int(q * len(per_frame)))]  # noqa: E731
        out["frame_spread"] = {"p10": pct(0.10), "p50": pct(0.50), "p90": pct(0.90)}
        out["stats"] = dev.stats()
        t = torch.tensor([ms_total], device=dev_s)
        if N > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_step"] = float(t.item()) / steps
        out["clocks"] = clocks.stop() if rank == 0 else None

        # ---- the same loop sustained for >= 1 s: the burst above runs at boost clock, an issue-bound kernel follows the SM clock ----
        if sustain_s > 0:
            n_s = max(steps, int(sustain_s * 1e3 / max(out["ms_step"], 1e-3)) + 1)
            sclk = ClockSampler(local_rank)
            if rank == 0:
                sclk.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(n_s):
                step()
            side_join()
            e1.record(stream)
            torch.cuda.synchronize()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev_s)
            if N > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ck = sclk.stop() if rank == 0 else None
            out["sustained"] = {"ms_per_step": float(t.item()) / n_s, "steps": n_s, "seconds": float(t.item()) * 1e-3, "clocks": ck}

        # ---- diagnostics (SWCU_TIMELINE=<frames>): where every kernel of a few pipelined frames really ran, per rank, written to
        #      gpurun_out/timeline_<workload>_n<N>_r<rank>.json (times in ms since the first kernel of the rank's timeline) ----
        tl_frames = int(os.environ.get("SWCU_TIMELINE", "0"))
        if tl_frames > 0:
            barrier()
            dev.set_profiling(2)
            for _ in range(tl_frames):
                step()
            tl = dev.timeline()
            dev.set_profiling(0)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"timeline_{name}_n{N}_r{rank}.json"), "w") as f:
                json.dump(tl, f)
            barrier()

        # ---- end to end through the C-ABI with HOST buffers: H2D of the step's inputs and D2H of the frame inside the timed region.
        #      A render loop with several frames in flight, bounded by a fence per frame (a Vulkan application's per-frame vk::Fence):
        #      the library's copy streams bring frame i+1's inputs up and frame i-1's pixels down while frame i renders.  At N > 1
        #      every rank uploads 1/N of each input buffer over its own PCIe link and the slices are all-gathered over NVLink (NCCL);
        #      rank 0 downloads the assembled frame from the ring ----
        if with_e2e:
            e2e_steps = max(6, min(steps, 30))
            alt_inputs, alt_keep = [], []
            alt_descs = [sc.build_desc(dataclasses.replace(dr, vertices=np.array(dr.vertices, dtype=np.float32, copy=True),
                                                           indices=None if dr.indices is None else dr.indices.copy()),
                                       frame.att, alt_keep, area, alt_inputs) for dr in sc.draws]
            for b in alt_inputs:
                dev.register(b, upload=False)
            in_sets = [(frame.inputs, frame.descs), (alt_inputs, alt_descs)]
            sharded_upload = N > 1 and os.environ.get("SWCU_SHARD_UPLOAD", "1") != "0"
            gather_stream = torch.cuda.Stream() if sharded_upload else None
            gather_views = {}
            up_bytes = [0]

            def upload_set(k):
                for b in in_sets[k][0]:
                    if not sharded_upload:
                        dev.upload(b)
                        up_bytes[0] += b.nbytes
                        continue
                    # my 1/N over PCIe, the rest over NVLink: all-gather of the 16-byte-aligned bulk, every rank uploads the small tail
                    flat = b.reshape(-1).view(np.uint8)
                    chunk = (flat.nbytes // N) & ~15
                    tail = flat.nbytes - chunk * N
                    if chunk:
                        dev.check(dev.lib.swcu_mem_upload(dev.ctx, flat.ctypes.data + rank * chunk, chunk))
                    if tail:
                        dev.check(dev.lib.swcu_mem_upload(dev.ctx, flat.ctypes.data + chunk * N, tail))
                    if chunk:
                        # the collective runs on a stream of its own, beside the frame that is rendering: it waits for my slice's upload,
                        # the draw that reads the buffer waits for it
                        dev.check(dev.lib.swcu_mem_acquire_on(dev.ctx, flat.ctypes.data, gather_stream.cuda_stream))
                        key = (b.ctypes.data, chunk)
                        if key not in gather_views:  # (wrapping the shadow in a tensor costs more host time than the collective's launch)
                            gather_views[key] = torch.as_tensor(_DevArr(dev.device_ptr(b), chunk * N), device=dev_s)
                        whole = gather_views[key]
                        with torch.cuda.stream(gather_stream):
                            dist.all_gather_into_tensor(whole, whole[rank * chunk:(rank + 1) * chunk])
                        dev.check(dev.lib.swcu_mem_release_on(dev.ctx, flat.ctypes.data, gather_stream.cuda_stream))
                    up_bytes[0] += chunk + tail

            F = len(final_imgs) if (rank == 0 or N == 1) else 2
            state["i"] = 0
            if pg is not None:  # the e2e loop starts on a slot boundary of the ring
                while pg.frame_no % pg.slots:
                    step()
                dev.sync()
                barrier()

            def e2e_frame(i):
                upload_set((i + 1) % 2)  # next frame's inputs, behind this frame's on the upload stream
                if F == 1 and i >= 1 and rank == 0:
                    dev.fence_wait((i - 1) % 2)  # the one host image: frame i-1 is consumed before frame i may land in it
                step(present=lambda img: dev.download(img), descs=in_sets[i % 2][1])
                dev.fence_signal(i % max(F, 2))
                if F > 1 and i >= F - 1:
                    dev.fence_wait((i - (F - 1)) % F)  # frame i - (F - 1) is on the host now

            def e2e_run(n):
                upload_set(0)  # frame 0's inputs; every frame of the loop uploads one set, so n frames move n sets
                for i in range(n):
                    e2e_frame(i)
                dev.sync()

            e2e_run(3)
            barrier()
            torch.cuda.synchronize()
            up_bytes[0] = 0
            t0 = time.perf_counter()
            e2e_run(e2e_steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device=dev_s)
            if N > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["e2e"] = {"ms_per_step": float(t.item()) * 1e3 / e2e_steps, "steps": e2e_steps, "frames_in_flight": max(F, 2),
                          "h2d_bytes_per_step": int(up_bytes[0] / (e2e_steps + 1)) * N,  # all ranks together
                          "h2d_bytes_per_step_per_gpu": int(up_bytes[0] / (e2e_steps + 1)),
                          "d2h_bytes_per_step": H * pitch,
                          "inputs": "1/N per rank over PCIe + NCCL all-gather over NVLink" if sharded_upload else "every rank uploads everything" if N > 1 else "uploaded"}

        # ---- per-kernel device times (events around every launch) for the roofline of the dominant kernel ----
        dev.set_profiling(True)
        per = {}
        for _ in range(5):
            for d_ in frame.descs:
                dev.draw(d_)
                for kname, ms in dev.last_draw_kernels():
                    per.setdefault(kname, []).append(ms)
        dev.set_profiling(False)
        dev.sync()
        torch.cuda.synchronize()
        barrier()
    out["kernels"] = {k: statistics.mean(v) for k, v in per.items()}
    if pg is not None:
        pg.close()
    frame.close()
    for f in finals:
        if f is not frame.resolved:
            dev.unregister(f)
    if group is not None:
        group.close()
    dev.close()
    return out


def workload_line(name: str, r: dict, N: int, steps: int, warmup: int) -> dict:
    """The JSON pieces of one measured workload."""
    wl = r["wl"]
    sc = wl.scene
    ms_step = r["ms_step"]
    peak, peak_src = measured_peaks()
    tile_name = "k_tile<4>" if sc.samples == 4 else "k_tile<1>"
    tile_ms = r["kernels"].get(tile_name)
    tile_bytes = wl.tile_bytes / N
    achieved = tile_bytes / (tile_ms * 1e-3) / 1e9 if tile_ms else None
    frame_bytes = wl.algorithmic_bytes if N == 1 else None
    # DRAM bytes of the dominant kernel per launch, from the committed ncu capture of this workload — only while the kernel sources
    # are the ones the capture was taken from (profiles/traffic.json carries their hash)
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if N == 1 and wl.name in tj and tj[wl.name]["kernel"] == tile_name:
            if tj[wl.name].get("kernels_sha16") == kernels_hash():
                traffic = tj[wl.name]["bytes"]
                traffic_note = tj[wl.name].get("source")
            else:
                traffic_note = "stale: the kernel sources changed after the committed capture"
    except (OSError, ValueError, KeyError):
        pass
    line = {
        "value": wl.covered_pixels / (ms_step * 1e-3) / 1e9, "unit": "Gpixels/s", "ms_per_step": ms_step, "ms_per_step_spread": r["frame_spread"],
        "mtris_per_s": wl.triangles / (ms_step * 1e-3) / 1e6, "host_enqueue_ms_per_step": r.get("host_enqueue_ms_step"),
        "frame_hash_ok": r.get("frame_hash_ok"),
        "config": {"workload": wl.name, "description": wl.description, "bands": N, "band_rows": r["band_rows"],
                   "l2": "inputs larger than L2 (framebuffer + mesh + per-triangle records > 126 MB)" if wl.algorithmic_bytes > 200e6 else "working set fits L2; steady-state frames",
                   "step": "draw (+ resolve)" + ("" if N == 1 else (" with the setup sharded by triangle range (records stored into the owning rank's buffers over NVLink)" if r["group"] else " with replicated setup")
                                                 + (" + bands stored into rank 0's frame over NVLink (CUDA IPC) + flags" if r["gather"] == "peer" else " + NCCL all-gather of bands")),
                   "gather": r["gather"], "group": r["group"]},
        "clocks": r["clocks"],
        "gpu_launches": int(r["stats"].kernelLaunches),
        "gpu_launches_note": "kernels of libswcuda.so launched in the timed region (every kernel of the draw path is the library's own)",
        "kernels_ms": r["kernels"],
        "roofline": {"bound": "hbm", "kernel": tile_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": tile_bytes,
                     "frame_algorithmic_bytes": frame_bytes,
                     "frame_achieved": (frame_bytes / (ms_step * 1e-3) / 1e9) if frame_bytes else None,
                     "frame_frac": (frame_bytes / (ms_step * 1e-3) / 1e9 / peak) if frame_bytes else None},
    }
    if "sustained" in r:
        s = r["sustained"]
        line["sustained"] = {"value": wl.covered_pixels / (s["ms_per_step"] * 1e-3) / 1e9, "unit": "Gpixels/s", "ms_per_step": s["ms_per_step"],
                             "steps": s["steps"], "seconds": s["seconds"], "clocks": s["clocks"]}
    if "e2e" in r:
        e = r["e2e"]
        line["e2e"] = {"value": wl.covered_pixels / (e["ms_per_step"] * 1e-3) / 1e9, "unit": "Gpixels/s", **e}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 20
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if args.steps is None:  # the 1080p-4K single triangles are launch-bound -> steady state over many back-to-back frames
        args.steps = 256 if args.workload in ("c1", "c2", "c3") else 50

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the draw path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    N = world
    stream = torch.cuda.Stream()

    primary = run_workload(args.workload, args.steps, args.warmup, rank, N, local_rank, stream)
    secondary = None
    sec_name = None if args.no_secondary else SECONDARY.get(args.workload)
    if sec_name:
        sec_steps = max(5, min(args.steps, 20))
        secondary = run_workload(sec_name, sec_steps, args.warmup, rank, N, local_rank, stream, sustain_s=0.0)

    if rank == 0:
        wl = primary["wl"]
        body = workload_line(args.workload, primary, N, args.steps, args.warmup)
        line = {"metric": METRIC, "value": body.pop("value"), "unit": body.pop("unit"), "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": body.pop("ms_per_step"), "ms_per_step_spread": body.pop("ms_per_step_spread"), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32+u8 (fp32 setup/interpolation/blend, 24.8 fixed-point coverage, u16 sampler)", "data": "synthetic"}
        line.update(body)
        if secondary is not None:
            sb = workload_line(sec_name, secondary, N, sec_steps, args.warmup)
            sb["steps"] = sec_steps
            line["secondary"] = sb
        if N == 1 and not args.no_cpu_baseline:
            try:
                frames = max(3, min(40, int(15000 / EST_REF_MS[args.workload])))
                tm = reference_time(wl, frames, 1)
                line["cpu_baseline"] = {"value": wl.covered_pixels / (tm["median_ms"] * 1e-3) / 1e9, "unit": "Gpixels/s", "cores": host_threads(),
                                        "kind": "reference", "ms_per_step": tm["median_ms"],
                                        "sample": f"{frames} whole frames of {wl.name} on the unmodified reference ICD (oracle/_ref), draw-only LOAD pass + resolve, median"}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "Gpixels/s", "cores": host_threads(), "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    if N > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
