#!/usr/bin/env python
"""bench.py — the draw hot path on N B200s, one JSON line (see the contract in the task statement / DESIGN.md §Measurement).

A "step" is one frame of the workload: every draw of the scene through swcu_draw (setup -> spans -> binning -> tile
raster/shade/blend) plus, for the multisampled workload, the end-of-pass resolve; at N > 1 each rank renders its
screen band (renderArea = band) and the finished bands are all-gathered over NCCL.  Attachments stay resident; the clear
is outside the timed region, like the reference harness' LOAD pass (oracle/refrender.cpp --time).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl cuda|reference]
"""
from __future__ import annotations

import argparse
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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def run_reference(args, rank: int):
    if rank != 0:
        return
    from swiftshader_b200 import workloads
    wl = workloads.WORKLOADS[args.workload]()
    # bounded sample: the reference renders the whole frame K times; K is clamped so the run ends within minutes
    est_ms = {"c1": 3, "c2": 4, "c3": 15, "c4": 700, "c5": 4000}[args.workload]
    frames = max(1, min(args.steps, int(120000 / est_ms)))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    import numpy as np
    import torch
    import torch.distributed as dist
    from swiftshader_b200 import bands, workloads
    from swiftshader_b200.scene import Device, Frame
    from swiftshader_b200 import capi
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the draw path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    N = world

    wl = workloads.WORKLOADS[args.workload]()
    sc = wl.scene
    H, W = sc.height, sc.width
    band = bands.band_rows(H, N, rank)
    area = bands.render_area(W, H, N, rank)

    dev = Device(local_rank)
    stream = torch.cuda.Stream()
    dev.set_stream(stream.cuda_stream)
    frame = Frame(dev, sc, render_area=area)
    frame.upload_inputs()
    frame.clear()
    H2 = sc.padded_height()
    pitch = W * 4

    # resolve only my band: attachments re-based to the band's first row
    def band_att(host_arr, slice_b):
        return capi.Attachment(host_arr.ctypes.data + band[0] * pitch, sc.colorFormat, pitch, slice_b, W, band[1] - band[0], 0)

    src_b = band_att(frame.att["color"], H2 * pitch)
    dst_b = band_att(frame.resolved, H2 * pitch) if frame.resolved is not None else None

    class _DevArr:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    # N > 1: how the finished bands reach rank 0.  "peer" (default): the end-of-pass resolve (4x) / a band copy (1x) of every
    # other rank stores straight into rank 0's frame over NVLink (CUDA IPC mapping), ordered by flags — no collective on the
    # data path.  "nccl": in-place NCCL all-gather of the bands (SWCU_GATHER=nccl).
    gather = os.environ.get("SWCU_GATHER", "peer") if N > 1 else "none"
    full = None
    pg = None
    if gather == "nccl":
        full = torch.as_tensor(_DevArr(frame.final_device_ptr(), H * pitch), device=f"cuda:{local_rank}")
    elif gather == "peer":
        try:
            pg = bands.PeerGather(dev, frame.final_image(), H, pitch, N, rank)
            ok = 1
        except Exception as e:  # noqa: BLE001  (e.g. CUDA IPC not permitted in this container)
            print(f"[bench] rank {rank}: peer delivery unavailable ({e}); falling back to the NCCL all-gather", file=sys.stderr)
            pg, ok = None, 0
        t_ok = torch.tensor([ok], device=f"cuda:{local_rank}")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            pg, gather = None, "nccl"
            full = torch.as_tensor(_DevArr(frame.final_device_ptr(), H * pitch), device=f"cuda:{local_rank}")
        elif rank != 0:
            peer_dst = pg.band_destination(sc.colorFormat, W)

    def step(present=None, descs=None):
        for d_ in (descs if descs is not None else frame.descs):
            dev.draw(d_)
        if pg is not None and rank != 0:
            pg.begin_frame()  # rank 0 must be done with the previous frame before its rows are overwritten
            if dst_b is not None:
                dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src_b), sc.samples, C.byref(peer_dst)))
            else:
                dev.check(dev.lib.swcu_copy_image(dev.ctx, C.byref(src_b), C.byref(peer_dst)))
            pg.band_done()
            return
        if dst_b is not None:
            dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src_b), sc.samples, C.byref(dst_b)))
        if pg is not None:
            pg.begin_frame()
            pg.band_done()  # the stream waits for the other ranks' bands
            if present is not None:
                present()
            pg.frame_consumed()
            return
        if gather == "nccl":
            bands.gather_bands(full, H, pitch, N, rank)  # NCCL all-gather, in place: my band is already at its slot
        if present is not None:
            present()

    def barrier():
        if N > 1:
            dist.barrier()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        dev.reset_stats()
        # one event per frame boundary: the contract's value is total / K; the per-frame spread (p10 / p50 / p90) is reported too
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        torch.cuda.synchronize()
        marks[0].record(stream)
        for i in range(args.steps):
            step()
            marks[i + 1].record(stream)
        torch.cuda.synchronize()
        barrier()
        ms_total = marks[0].elapsed_time(marks[-1])
        per_frame = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))
        pct = lambda q: per_frame[min(len(per_frame) - 1, int(q * len(per_frame)))]  # noqa: E731
        frame_spread = {"p10": pct(0.10), "p50": pct(0.50), "p90": pct(0.90)}
        st = dev.stats()
        clock_info = clocks.stop() if rank == 0 else None
        t = torch.tensor([ms_total], device=f"cuda:{local_rank}")
        if N > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item()) / args.steps

        # ---- end to end through the C-ABI with HOST buffers: H2D of the step's inputs and D2H of the frame inside the timed region.
        #      A render loop with several frames in flight, bounded by a fence per frame (a Vulkan application's per-frame vk::Fence):
        #      the library's copy streams bring frame i+1's inputs up and frame i-1's pixels down while frame i renders ----
        e2e_steps = max(6, min(args.steps, 30))
        if gather == "nccl":
            dev.set_option("copy_streams", 0)  # the all-gather writes the frame on the caller's stream, unseen by the library
        # Three stages overlap - frame i+1's inputs going up, frame i rendering, frame i-1 coming down - so the loop holds two
        # sets of input buffers (per-frame dynamic vertex data, as an application double-buffers it) and, at N = 1 with MSAA,
        # three resolve targets; a fence per frame bounds the frames in flight.
        import dataclasses
        alt_inputs, alt_keep = [], []
        alt_descs = [sc.build_desc(dataclasses.replace(dr, vertices=np.array(dr.vertices, dtype=np.float32, copy=True),
                                                       indices=None if dr.indices is None else dr.indices.copy()),
                                   frame.att, alt_keep, area, alt_inputs) for dr in sc.draws]
        for b in alt_inputs:
            dev.register(b, upload=False)
        in_sets = [(frame.inputs, frame.descs), (alt_inputs, alt_descs)]
        skip = set(filter(None, os.environ.get("SWCU_E2E_SKIP", "").split(",")))  # diagnosis only: the reported e2e runs with nothing skipped

        def upload_set(k):
            for b in in_sets[k][0]:
                dev.upload(b)

        if N == 1:
            outs, dsts = [frame.final_image()], [dst_b]
            if dst_b is not None:
                for _ in range(2):
                    extra = np.zeros_like(frame.resolved)
                    dev.register(extra, upload=False)
                    outs.append(extra[0])
                    dsts.append(band_att(extra, H2 * pitch))
            F = len(outs)  # frames in flight (1x: the colour attachment itself is the only host-visible image)

            def e2e_frame(i):
                upload_set((i + 1) % 2)  # next frame's inputs, behind this frame's on the upload stream
                for d_ in in_sets[i % 2][1]:
                    dev.draw(d_)
                k = i % F
                if dsts[k] is not None:
                    dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src_b), sc.samples, C.byref(dsts[k])))
                if F == 1 and i >= 1:
                    dev.fence_wait((i - 1) % 2)  # the one host image: frame i-1 is consumed before frame i may land in it
                if "download" not in skip:
                    dev.download(outs[k])
                dev.fence_signal(i % max(F, 2))
                if F > 1 and i >= F - 1:
                    dev.fence_wait((i - (F - 1)) % F)  # frame i-2 is on the host now
            frames_in_flight = max(F, 2)
        else:
            # rank 0's frame is the one host-visible image (the other ranks store their bands into it): frame i-1 is consumed
            # before frame i comes down; the consumed flag leaves rank 0 from its download stream, so its next draw is not held back
            def present(i):
                if i >= 1:
                    dev.fence_wait((i - 1) % 2)
                frame.download_final()

            def e2e_frame(i):
                upload_set((i + 1) % 2)
                step((lambda: present(i)) if rank == 0 else None, in_sets[i % 2][1])
                dev.fence_signal(i % 2)
                if rank != 0 and i >= 1:
                    dev.fence_wait((i - 1) % 2)
            frames_in_flight = 2

        def e2e_run(n):
            upload_set(0)  # frame 0's inputs; every frame of the loop uploads one set, so n frames move n sets
            for i in range(n):
                e2e_frame(i)
            dev.sync()

        e2e_run(3)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=f"cuda:{local_rank}")
        if N > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) * 1e3 / e2e_steps
        h2d = frame.input_bytes()
        d2h = H * pitch

        # ---- per-kernel device times (events around every launch) for the roofline of the dominant kernel ----
        dev.set_profiling(True)
        per = {}
        prof_steps = 5
        for _ in range(prof_steps):
            frame.draw()
            for name, ms in dev.last_draw_kernels():
                per.setdefault(name, []).append(ms)
        dev.set_profiling(False)
        torch.cuda.synchronize()

    kernels = {k: statistics.mean(v) for k, v in per.items()}
    peak, peak_src = measured_peaks()
    tile_name = "k_tile<4>" if sc.samples == 4 else "k_tile<1>"
    tile_ms = kernels.get(tile_name)
    tile_bytes = wl.tile_bytes / N
    achieved = tile_bytes / (tile_ms * 1e-3) / 1e9 if tile_ms else None
    frame_bytes = wl.algorithmic_bytes if N == 1 else None
    traffic = None  # DRAM bytes of the dominant kernel per launch, from the committed ncu capture of this workload (if any)
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")))
        if N == 1 and wl.name in tj and tj[wl.name]["kernel"] == tile_name:
            traffic = tj[wl.name]["bytes"]
    except (OSError, ValueError, KeyError):
        pass

    if rank == 0:
        gpix = wl.covered_pixels / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": gpix, "unit": "Gpixels/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "ms_per_step_spread": frame_spread, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8 (fp32 setup/interpolation/blend, 24.8 fixed-point coverage, u16 sampler)",
            "data": "synthetic",
            "mtris_per_s": wl.triangles / (ms_step * 1e-3) / 1e6,
            "config": {"workload": wl.name, "description": wl.description, "bands": N, "band_rows": band[1] - band[0],
                       "l2": "inputs larger than L2 (framebuffer + mesh + per-triangle records > 126 MB)" if wl.algorithmic_bytes > 200e6 else "working set fits L2; steady-state frames",
                       "step": "draw (+ resolve)" + ("" if N == 1 else " + bands stored into rank 0's frame over NVLink (CUDA IPC) + flags" if gather == "peer" else " + NCCL all-gather of bands"),
                       "gather": gather},
            "clocks": clock_info,
            "e2e": {"value": wl.covered_pixels / (e2e_ms * 1e-3) / 1e9, "unit": "Gpixels/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "frames_in_flight": frames_in_flight},
            "gpu_launches": int(st.kernelLaunches),
            "gpu_launches_note": "kernels of libswcuda.so launched in the timed region (excludes the CUB scan/sort kernels between them)",
            "kernels_ms": kernels,
            "roofline": {"bound": "hbm", "kernel": tile_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": tile_bytes,
                         "frame_algorithmic_bytes": frame_bytes,
                         "frame_achieved": (frame_bytes / (ms_step * 1e-3) / 1e9) if frame_bytes else None,
                         "frame_frac": (frame_bytes / (ms_step * 1e-3) / 1e9 / peak) if frame_bytes else None},
        }
        if N == 1 and not args.no_cpu_baseline:
            try:
                est_ms = {"c1": 3, "c2": 4, "c3": 15, "c4": 700, "c5": 4000}[args.workload]
                frames = max(3, min(40, int(15000 / est_ms)))
                tm = reference_time(wl, frames, 1)
                line["cpu_baseline"] = {"value": wl.covered_pixels / (tm["median_ms"] * 1e-3) / 1e9, "unit": "Gpixels/s", "cores": host_threads(),
                                        "kind": "reference", "ms_per_step": tm["median_ms"],
                                        "sample": f"{frames} whole frames of {wl.name} on the unmodified reference ICD (oracle/_ref), draw-only LOAD pass + resolve, median"}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "Gpixels/s", "cores": host_threads(), "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    if pg is not None:
        pg.close()
    frame.close()
    dev.close()
    if N > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
