"""Developer check: pinned host<->device copy bandwidth on this box (what bounds bench.py's e2e)."""
import time, torch
torch.cuda.init()
for gb in (1, 8):
    n = gb * (1 << 30)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        print(gb, "GiB", name, "%.1f GB/s" % (3 * n / (time.perf_counter() - t) / 1e9))
    # both directions at once on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty(n // 8, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    with torch.cuda.stream(s1):
        for _ in range(3):
            h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        for _ in range(3):
            d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    print(gb, "GiB duplex d2h(+h2d/8)", "%.1f GB/s d2h" % (3 * n / (time.perf_counter() - t) / 1e9))
    del h, d, h2, d2
