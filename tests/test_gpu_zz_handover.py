"""GPU: the hand-over stream of the C-ABI (swcu_side_begin / swcu_side_end / swcu_side_wait, include/swcu.h) in the shape the
multi-GPU frame loop uses it (bench.py): the main stream copies a finished frame into a band buffer and goes on with the next
frame, the delivery of the buffer — a second copy and the download — runs beside it on the hand-over stream; the buffer is only
written again after swcu_side_wait.  Both delivered frames must be the oracle's, bit for bit.
(Also runnable without pytest / torch: python tests/test_gpu_zz_handover.py)"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import scenes  # noqa: E402
from oracle import swref  # noqa: E402


def handover_check(device):
    from swiftshader_b200.scene import Frame
    sc = scenes.blend(3)  # one blended draw, 1x BGRA8: a second pass over the same attachments changes the frame
    H = sc.height
    want1 = swref.render_oracle(sc)["color"].copy()
    att = swref.render_oracle(sc)
    want2 = swref.render_oracle(sc, att=att)["color"].copy()
    assert not np.array_equal(want1, want2)
    lib, ctx = device.lib, device.ctx
    fr = Frame(device, sc)
    mid, out1, out2 = (np.zeros_like(fr.att["color"]) for _ in range(3))
    for b in (mid, out1, out2):
        device.register(b, upload=False)
    try:
        def att_of(arr):
            a = fr._attachment("color")
            a.buffer = arr.ctypes.data
            return a
        color, a_mid, a1, a2 = fr._attachment("color"), att_of(mid), att_of(out1), att_of(out2)
        fr.upload_inputs()
        fr.upload_attachments()
        fr.draw()                                                                    # frame 1
        device.check(lib.swcu_copy_image(ctx, C.byref(color), C.byref(a_mid)))       # main stream: frame 1 -> band buffer
        device.check(lib.swcu_side_begin(ctx))
        device.check(lib.swcu_copy_image(ctx, C.byref(a_mid), C.byref(a1)))          # hand-over stream: delivery + download
        device.download(out1)
        device.check(lib.swcu_side_end(ctx, 0))
        fr.draw()                                                                    # frame 2, beside the delivery of frame 1
        device.check(lib.swcu_side_wait(ctx, 0))                                     # the band buffer has been read
        device.check(lib.swcu_copy_image(ctx, C.byref(color), C.byref(a_mid)))
        device.check(lib.swcu_side_begin(ctx))
        device.check(lib.swcu_copy_image(ctx, C.byref(a_mid), C.byref(a2)))
        device.download(out2)
        device.check(lib.swcu_side_end(ctx, 1))
        device.sync()
        assert np.array_equal(out1[0][:H], want1[0][:H]), "frame 1 as delivered by the hand-over stream differs from the oracle"
        assert np.array_equal(out2[0][:H], want2[0][:H]), "frame 2 as delivered by the hand-over stream differs from the oracle"
        # a second swcu_side_begin without an end, and an end without a begin, are refused
        device.check(lib.swcu_side_begin(ctx))
        assert lib.swcu_side_begin(ctx) != 0
        device.check(lib.swcu_side_end(ctx, 0))
        assert lib.swcu_side_end(ctx, 0) != 0
    finally:
        device.sync()
        for b in (mid, out1, out2):
            device.unregister(b)
        fr.close()


@pytest.mark.gpu
def test_handover_stream_delivers_frames_beside_the_next_one(device):
    handover_check(device)


if __name__ == "__main__":
    from swiftshader_b200.scene import Device
    dev = Device(0)
    try:
        handover_check(dev)
        print("handover ok")
    finally:
        dev.close()
