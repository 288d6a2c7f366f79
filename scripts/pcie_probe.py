"""Probe of the host link: H2D and D2H alone and together (two streams), pinned memory.  Prints GB/s."""
import time
import torch

dev = torch.device("cuda:0")
nh, nd = 26 << 20, 33 << 20
h_in = torch.empty(nh, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nd, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nh, dtype=torch.uint8, device=dev)
d_out = torch.empty(nd, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for _ in range(2):
    run(True, True, 3)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D {nh / a / 1e6:.1f} GB/s ({a:.3f} ms)  D2H {nd / b / 1e6:.1f} GB/s ({b:.3f} ms)  both {c:.3f} ms (sum {a + b:.3f}, max {max(a, b):.3f})")
