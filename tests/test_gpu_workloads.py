"""GPU: the five BASELINE.json workloads at FULL size against the CPU oracle (bit-exact), band sharding (what each rank
of a multi-GPU run renders) against the single-band image, and the device-side clear / resolve either side of the draw."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenes
from oracle import swref
from swiftshader_b200 import workloads
from swiftshader_b200.scene import Frame

pytestmark = pytest.mark.gpu

# sha256 of what the UNMODIFIED reference ICD renders for each workload at full size (tests/golden/gen_workload_hashes.py)
REF_HASHES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "workload_hashes.json")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check_reference_hashes(key, scene, got):
    """Colour (the resolved 1x image when multisampled), depth and stencil (single-sampled) against the reference ICD's own render."""
    want = REF_HASHES[key]["hashes"]
    H = scene.height
    img = got["resolved"][0, :H] if scene.samples > 1 else got["color"][0, :H]
    assert _sha(img) == want["color"], f"{key}: colour differs from the reference ICD's render"
    if scene.samples == 1:
        for k in ("depth", "stencil"):
            if k in want:
                assert _sha(got[k][0, :H]) == want[k], f"{key}: {k} differs from the reference ICD's render"


def _render_frame(device, scene, area=None):
    fr = Frame(device, scene, render_area=area)
    try:
        fr.upload_inputs()
        fr.clear()
        fr.draw()
        fr.resolve()
        fr.download_all()
        device.sync()
        out = {k: v.copy() for k, v in fr.att.items()}
        if fr.resolved is not None:
            out["resolved"] = fr.resolved.copy()
        return out
    finally:
        fr.close()


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4"])
def test_full_size_workload_bit_exact_vs_oracle(device, name):
    wl = workloads.WORKLOADS[name]()
    got = _render_frame(device, wl.scene)
    _check_reference_hashes(name, wl.scene, got)
    want = swref.render_oracle(wl.scene)
    for k in want:
        assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), f"{name}/{k}: {(got[k] != want[k]).sum()} elements differ"
    if wl.scene.samples > 1:
        res = swref.resolve_oracle(wl.scene, want)
        assert np.array_equal(got["resolved"][0], res)
    # every pixel of the half-screen triangles / full-screen meshes is shaded exactly once per layer
    if name in ("c1", "c2", "c3"):
        clear = wl.scene.clear_color_bytes()
        touched = (got["color"][0, :wl.scene.height] != clear).any(axis=-1).sum()
        assert touched == wl.covered_pixels


def test_c5_full_size_properties(device):
    """10M triangles at 8K: (1) a single-layer mesh covers every pixel exactly once -> no clear colour left and the
    image is idempotent under a second identical frame (no blending); (2) 48 rows spread over the frame equal the oracle's
    render of the same rows (scissored), bit-exact."""
    wl = workloads.c5()
    sc = wl.scene
    fr = Frame(device, sc)
    try:
        fr.upload_inputs(); fr.clear(); fr.draw(); fr.download_all(); device.sync()
        first = fr.att["color"].copy()
        # all 4320 rows against the reference ICD's own render of the 10 M triangles
        _check_reference_hashes("c5", sc, {"color": first})
        fr.draw(); fr.download_all(); device.sync()
        assert np.array_equal(first, fr.att["color"])
    finally:
        fr.close()
    clear = sc.clear_color_bytes()
    # alpha of the noise texture is random, so compare all four channels against the clear colour
    assert ((first[0, :sc.height] != clear).any(axis=-1)).mean() > 0.99
    for y0 in (0, 2000, sc.height - 16):
        sc.draws[0].scissor = (0, y0, sc.width, 16)
        want = swref.render_oracle(sc)
        sc.draws[0].scissor = None
        assert np.array_equal(first[0, y0:y0 + 16], want["color"][0, y0:y0 + 16]), f"rows {y0}..{y0 + 16}"


@pytest.mark.parametrize("name", ["c4", "c5", "c2"])
@pytest.mark.parametrize("nbands", [2, 4, 8])
def test_bands_reassemble_to_the_single_gpu_image(device, name, nbands):
    """Each rank renders renderArea = its band; the union of the bands must equal the 1-band frame bit for bit."""
    wl = workloads.small(name)
    sc = wl.scene
    H = sc.height
    if H % (2 * nbands):
        pytest.skip("height does not split into even bands")
    whole = _render_frame(device, sc)
    for k in whole:
        merged = np.zeros_like(whole[k])
        for r in range(nbands):
            y0, y1 = r * H // nbands, (r + 1) * H // nbands
            part = _render_frame(device, sc, area=(0, y0, sc.width, y1 - y0))
            merged[:, y0:y1] = part[k][:, y0:y1]
        assert np.array_equal(merged[:, :H].view(np.uint8), whole[k][:, :H].view(np.uint8)), f"{name}/{k} with {nbands} bands"


def test_device_clear_and_resolve_match_oracle(device):
    sc = scenes.msaa(3)
    got = _render_frame(device, sc)
    want = swref.render_oracle(sc)  # host-side clear of alloc_attachments + oracle draw
    for k in want:
        assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8))
    assert np.array_equal(got["resolved"][0], swref.resolve_oracle(sc, want))


def test_d16_device_clear_and_unaligned_pitch(device):
    """D16_UNORM: the device-side clear (2-byte fill) and a framebuffer whose row pitch does not satisfy the tensor-map rules
    (cooperative 16-bit tile copies instead of TMA), both against the oracle."""
    for sc in (scenes.depth16(0), scenes.depth16(8)):
        got = _render_frame(device, sc)
        want = swref.render_oracle(sc)
        for k in want:
            assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), k
    base = scenes.depth16(10)
    sc = scenes.Scene(100, 70, base.draws, hasDepth=True, depthFormat=scenes.FMT_D16_UNORM, clearDepth=1.0, clearColor=(0.1, 0.2, 0.3, 1.0))
    for binned in (0, 1):
        device.set_option("force_binned", binned)
        try:
            got = _render_frame(device, sc)
        finally:
            device.set_option("force_binned", 0)
        want = swref.render_oracle(sc)
        for k in want:
            assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), (k, binned)


def test_unsupported_state_is_a_hard_error(device):
    from swiftshader_b200 import capi
    sc = scenes.benchmark(1, 64, 64)
    sc.colorFormat = 64  # A2B10G10R10_UNORM_PACK32: outside the subset
    with pytest.raises(capi.SwcuError) as e:
        device.render(sc)
    assert e.value.code == capi.E_UNSUPPORTED


@pytest.mark.parametrize("samples", [1, 4])
def test_frames_in_flight_match_serial_frames(device, samples):
    """The copy streams and fences of the boundary (swcu_mem_upload / swcu_mem_download / swcu_fence_*): a render loop that
    keeps two frames in flight — new vertex data every frame, clear, draw, resolve, download — delivers exactly the frames
    of the same loop with a full swcu_sync after every call.  Covers an upload overlapping the previous frame's tile kernel,
    a download overlapping the next frame's clear / draw of the SAME attachment (1x) or the next resolve (4x)."""
    from swiftshader_b200.scene import Draw, Scene
    rng = np.random.default_rng(77)
    tris = [scenes._verts(rng, scenes._tri_kind(rng, (5, 1, 0)[i % 3]), persp=(i % 2 == 0), colour=rng.uniform(0, 1, (3, 4))) for i in range(400)]
    base = np.ascontiguousarray(np.concatenate(tris), dtype=np.float32)
    d = Draw(base.copy(), scenes.P4C4, "vs_pos4_col4", "fs_col4", depthTest=True, depthWrite=True, blend=True)
    sc = Scene(192, 160, [d], samples=samples, hasDepth=True, clearDepth=1.0, clearColor=(0.1, 0.2, 0.3, 1.0))
    frames = 7
    variants = []
    for i in range(frames):
        v = base.copy()
        v[:, 4:7] = np.float32(1.0) - v[:, 4:7] if i % 2 else v[:, 4:7] * np.float32(0.25 + 0.1 * i)
        v[:, 0] += np.float32(0.02 * i) * v[:, 3]
        variants.append(v)
    fr = Frame(device, sc)
    device.set_option("force_binned", 1)  # the setup phase of every draw is host-synchronised: the vertex array may be rewritten after draw()
    try:
        verts = fr.inputs[0]
        assert verts.shape == base.shape
        outs = [fr.final_image()]
        dsts = [fr._attachment("resolved")] if samples > 1 else [None]
        if samples > 1:
            second = np.zeros_like(fr.resolved)
            device.register(second, upload=False)
            outs.append(second[0])
            a = fr._attachment("resolved")
            a.buffer = second.ctypes.data
            dsts.append(a)
        import ctypes as C

        def issue(i, serial):
            drain = device.sync if serial else (lambda: None)
            verts[...] = variants[i]
            fr.upload_inputs()
            drain()
            fr.clear()
            drain()
            fr.draw()
            drain()
            k = i % len(outs)
            if dsts[k] is not None:
                src = fr._attachment("color")
                device.check(device.lib.swcu_resolve(device.ctx, C.byref(src), samples, C.byref(dsts[k])))
                drain()
            return k

        want = []
        for i in range(frames):
            k = issue(i, True)
            device.download(outs[k])
            device.sync()
            want.append(outs[k].copy())
        assert any(not np.array_equal(want[0], w) for w in want[1:])
        for o in outs:
            o[...] = 0
        got = [None] * frames
        for i in range(frames):
            k = issue(i, False)
            if len(outs) == 1 and i >= 1:  # one host image: frame i-1 is consumed before frame i may land in it
                device.fence_wait((i - 1) % 2)
                got[i - 1] = outs[0].copy()
            device.download(outs[k])
            device.fence_signal(i % 2)
            if len(outs) == 2 and i >= 1:
                device.fence_wait((i - 1) % 2)
                got[i - 1] = outs[(i - 1) % 2].copy()
        device.fence_wait((frames - 1) % 2)
        got[frames - 1] = outs[(frames - 1) % len(outs)].copy()
        for i in range(frames):
            assert np.array_equal(got[i], want[i]), f"frame {i} of the pipelined loop differs from the serial loop"
    finally:
        device.set_option("force_binned", 0)
        device.sync()
        if samples > 1:
            device.unregister(second)
        fr.close()


@pytest.mark.parametrize("binned", [0, 1])
@pytest.mark.parametrize("samples", [1, 4])
def test_sample_mask_without_an_enabled_sample_draws_nothing(device, samples, binned):
    """PixelRoutine.cpp:104-111: a sample mask with no enabled sample (bit 0 clear at one sample per pixel, Context.cpp:527) leaves
    every attachment untouched — swcu_draw returns before any kernel."""
    import dataclasses
    base = scenes.msaa(3) if samples == 4 else scenes.benchmark(1, 96, 64)
    draws = [dataclasses.replace(d, sampleMask=0xFFFFFFF0 if samples == 4 else 0xFFFFFFFE) for d in base.draws]
    sc = dataclasses.replace(base, draws=draws)
    device.set_option("force_binned", binned)
    try:
        got = device.render(sc)
    finally:
        device.set_option("force_binned", 0)
    want = sc.alloc_attachments()
    for k in want:
        assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), k
    want = swref.render_oracle(sc)
    for k in want:
        assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), k


@pytest.mark.parametrize("seed", [0, 4])
def test_index_range_outside_the_registered_buffer_is_rejected(device, seed):
    """The kernels do no bounds checks on the index stream, so a draw whose index range leaves the registered range is an
    error at the boundary (SWCU_E_INVALID), for lists (seed 0) and strips (seed 4); the same draw with the range inside is accepted."""
    from swiftshader_b200 import capi
    from swiftshader_b200.scene import Frame
    sc = scenes.topology(seed)
    assert sc.draws[0].indices is not None
    fr = Frame(device, sc)
    try:
        fr.upload_inputs(); fr.clear()
        good = fr.descs[0]
        device.draw(good)
        bad = capi.DrawDesc.from_buffer_copy(good)
        bad.primitiveCount = good.primitiveCount + 1
        with pytest.raises(capi.SwcuError) as e:
            device.draw(bad)
        assert e.value.code == capi.E_INVALID
        bad2 = capi.DrawDesc.from_buffer_copy(good)
        bad2.indexBuffer = good.indexBuffer + sc.draws[0].indices.nbytes  # starts one past the end
        with pytest.raises(capi.SwcuError) as e:
            device.draw(bad2)
        assert e.value.code == capi.E_INVALID
        device.sync()
    finally:
        fr.close()


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
@pytest.mark.parametrize("binned", [0, 1])
def test_reduced_workloads_match_the_reference_icd(device, name, binned):
    wl = workloads.small(name)
    device.set_option("force_binned", binned)
    try:
        got = _render_frame(device, wl.scene)
    finally:
        device.set_option("force_binned", 0)
    _check_reference_hashes(f"small_{name}", wl.scene, got)


def test_multi_gpu_bands_assemble_to_the_reference_frame():
    """N ranks (one process per GPU) render their bands, deliver them to rank 0 (stores over NVLink and the NCCL all-gather) and
    rank 0 compares the assembled frame with the reference ICD's hash and the oracle; skipped below two GPUs."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(here, "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MISMATCH" not in res.stdout
