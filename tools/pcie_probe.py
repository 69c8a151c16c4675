import torch, time
for mb in (1, 4, 8, 32):
    n = mb * 1024 * 1024
    d = torch.empty(n, dtype=torch.uint8, device='cuda'); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device='cuda'); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): h.copy_(d, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    d2h = 20 * n / (a.elapsed_time(b) * 1e-3) / 1e9
    a.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    h2d = 20 * n / (a.elapsed_time(b) * 1e-3) / 1e9
    t0 = time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{mb:3d} MB: D2H {d2h:.1f} GB/s, H2D {h2d:.1f} GB/s, duplex D2H+H2D each {20*n/dt/1e9:.1f} GB/s")
