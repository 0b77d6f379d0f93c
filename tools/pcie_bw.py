"""Raw pinned-memory device-to-host bandwidth of the box (with and without a concurrent host-to-device stream):
the ceiling of bench.py's `e2e` figure, which has to move 33.2 MB of decoded picture per 2160p picture.

    python tools/pcie_bw.py
"""
import torch, time
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, both in (("d2h only", False), ("d2h + h2d", True)):
    for _ in range(2):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    for _ in range(4):
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                for _ in range(1):
                    d2.copy_(h2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s: D2H %.1f GB/s" % (name, 4 * n / ms / 1e6))
