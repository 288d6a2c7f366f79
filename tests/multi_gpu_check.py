"""Run under torchrun on N GPUs: every rank renders its band of the reduced-size workloads through the CUDA path, the bands
are all-gathered over NCCL, and rank 0 compares the assembled frame with the CPU oracle's full-frame render (bit-exact).
  python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py"""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import swref  # noqa: E402
from swiftshader_b200 import bands, workloads  # noqa: E402
from swiftshader_b200.scene import Device, Frame  # noqa: E402


class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = Device(local)
    stream = torch.cuda.Stream()
    dev.set_stream(stream.cuda_stream)
    bad = 0
    import ctypes as C
    from swiftshader_b200 import capi
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import scenes
    wl_hashes = json.load(open(os.path.join(ROOT, "tests", "golden", "workload_hashes.json")))
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_hashes.json")))
    cases = [(n, workloads.small(n).scene, wl_hashes[f"small_{n}"]["hashes"]["color"], ("nccl", "peer", "group")) for n in ("c4", "c5", "c2")]
    # the general set-up kernel in a group (a transform in the vertex stage, lines, points: quads that are always clipped and
    # re-projected, delivered to the owners of their rows like any other record), instanced draws = several draws per frame
    for n, gen, seed in (("mvp_0", scenes.mvp, 0), ("mvp_4", scenes.mvp, 4), ("lines_1", scenes.lines, 1), ("lines_3", scenes.lines, 3),
                         ("points_0", scenes.points, 0), ("points_2", scenes.points, 2), ("instanced_0", scenes.instanced, 0)):
        cases.append((n, gen(seed), gold[n]["color"], ("group",)))
    for name, sc, want_sha, modes in cases:
        H, W = sc.height, sc.width
        if H % (2 * world):
            continue
        for mode in modes:
            # "group": the setup of every binned draw is sharded by triangle range as well (swcu_group_*: records, bin counts and big-list
            # entries stored into the owning rank's work buffers over NVLink); the bands are delivered like in "peer"
            grp = None
            if mode == "group":
                grp = bands.Group(dev, sc, world, rank, max(d.primitive_count() for d in sc.flat_draws()), 6, sc.samples)
                dev.set_option("force_binned", 1)  # the single triangles too: a share can be empty, a big triangle spans every band
            fr = Frame(dev, sc, render_area=bands.render_area(W, H, world, rank))
            y0, y1 = bands.band_rows(H, world, rank)
            pitch = W * 4
            with torch.cuda.stream(stream):
                fr.upload_inputs(); fr.clear()
                if mode == "nccl":
                    fr.draw(); fr.resolve()
                    full = torch.as_tensor(_DevArr(fr.final_device_ptr(), H * W * 4), device=f"cuda:{local}")
                    bands.gather_bands(full, H, W * 4, world, rank)
                    torch.cuda.synchronize()
                    got = full.cpu().numpy().reshape(H, W, 4)
                else:
                    # two frames, so the "previous frame consumed" handshake is exercised as well
                    pg = bands.PeerGather(dev, fr.final_image(), H, pitch, world, rank)
                    H2 = sc.padded_height()
                    for _ in range(3 if mode == "group" else 2):
                        if sc.samples > 1:
                            fr.clear()  # the blended 4x scene is not idempotent; its 4x attachments are private to the rank
                        fr.draw()
                        pg.begin_frame()
                        if rank == 0:
                            if sc.samples > 1:  # only MY band: the other rows of the frame belong to the other ranks' stores
                                src = capi.Attachment(fr.att["color"].ctypes.data + y0 * pitch, sc.colorFormat, pitch, H2 * pitch, W, y1 - y0, 0)
                                dst = capi.Attachment(fr.resolved.ctypes.data + y0 * pitch, sc.colorFormat, pitch, H2 * pitch, W, y1 - y0, 0)
                                dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src), sc.samples, C.byref(dst)))
                            pg.band_done()
                            fr.download_final()
                            pg.frame_consumed()
                        else:
                            dst = pg.band_destination(sc.colorFormat, W)
                            src = capi.Attachment(fr.att["color"].ctypes.data + y0 * pitch, sc.colorFormat, pitch, H2 * pitch, W, y1 - y0, 0)
                            if sc.samples > 1:
                                dev.check(dev.lib.swcu_resolve(dev.ctx, C.byref(src), sc.samples, C.byref(dst)))
                            else:
                                dev.check(dev.lib.swcu_copy_image(dev.ctx, C.byref(src), C.byref(dst)))
                            pg.band_done()
                    dev.sync()
                    torch.cuda.synchronize()
                    got = fr.final_image()[:H].copy()
                    pg.close()
            fr.close()
            if grp is not None:
                grp.close()
                dev.set_option("force_binned", 0)
            if rank == 0:
                want = swref.render_oracle(sc)
                ref = swref.resolve_oracle(sc, want) if sc.samples > 1 else want["color"][0]
                ok = np.array_equal(got, ref[:H])
                # ... and with what the unmodified reference ICD rendered for the same scene (tests/golden/gen_workload_hashes.py)
                ok = ok and hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == want_sha
                print(f"multi_gpu_check {name} world={world} gather={mode}: {'ok' if ok else 'MISMATCH'}", flush=True)
                bad += 0 if ok else 1
    dev.close()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
