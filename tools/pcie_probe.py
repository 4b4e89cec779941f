"""Ceiling of the end-to-end leg: pinned host <-> device copies of the config-2 step's bytes (2 x 2^24 x 40 B each way), one
direction alone and both directions concurrently on two streams (CUDA events, best of 5).  The host-pointer entry points
cannot beat the concurrent figure: every input byte crosses PCIe once and every output byte once."""
import torch

n = (1 << 24) * 2 * 40
dev = torch.device("cuda", 0)
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        e1.record(s1)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        e2.record(s2)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    return best


for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both directions", True, True)):
    ms = run(a, b)
    print(f"{name:16s} {ms:7.2f} ms for {n / 1e9:.2f} GB per direction = {n / ms / 1e6:.1f} GB/s per direction")
