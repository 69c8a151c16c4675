#!/usr/bin/env python
"""Host<->device copy bandwidth with N GPUs copying AT THE SAME TIME (one process per GPU, pinned buffers of 32 MB, copy engines only):
what bounds the value-array e2e number of bench.py at 2 / 4 / 8 GPUs (every rank moves 8 MB up and 31 MB down per step).
   gpurun --gpus 8 -- python tools/pcie_concurrent.py > gpurun_out/pcie_concurrent.md"""
import os
import subprocess
import sys
import time


def child(start_at, seconds):
    import torch
    n = 32 << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda"); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        h.copy_(d, non_blocking=True); d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    res = []
    for k, mode in enumerate(("d2h", "h2d", "duplex")):
        t_go = start_at + k * (seconds + 1.0)
        while time.time() < t_go:
            time.sleep(0.0005)
        t0 = time.perf_counter(); cnt = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(8):
                if mode in ("d2h", "duplex"):
                    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
                if mode in ("h2d", "duplex"):
                    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
            torch.cuda.synchronize(); cnt += 8
        dt = time.perf_counter() - t0
        res.append(cnt * n / dt / 1e9)
    print("%.2f %.2f %.2f" % tuple(res), flush=True)


def main():
    import torch
    ng = torch.cuda.device_count()
    print("| GPUs copying at once | D2H per GPU (GB/s) | D2H aggregate | H2D per GPU | H2D aggregate | duplex per GPU and direction | duplex aggregate (both directions) |")
    print("|---|---|---|---|---|---|---|")
    for n in [k for k in (1, 2, 4, 8) if k <= ng]:
        start = time.time() + 12.0 + 2.0 * n          # children import torch first
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "child", repr(start), "1.5"], stdout=subprocess.PIPE, text=True,
                                  env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(g))) for g in range(n)]
        rows = [[float(v) for v in p.communicate()[0].split()] for p in procs]
        d2h = [r[0] for r in rows]; h2d = [r[1] for r in rows]; dup = [r[2] for r in rows]
        print(f"| {n} | {min(d2h):.1f} – {max(d2h):.1f} | {sum(d2h):.0f} | {min(h2d):.1f} – {max(h2d):.1f} | {sum(h2d):.0f} | {min(dup):.1f} – {max(dup):.1f} | {2 * sum(dup):.0f} |", flush=True)
    print("\nnproc:", os.cpu_count(), " affinity:", len(os.sched_getaffinity(0)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(float(sys.argv[2]), float(sys.argv[3]))
    else:
        main()
