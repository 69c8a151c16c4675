#!/usr/bin/env python
"""Per-source-line stall/instruction summary of an Nsight Compute report (needs -lineinfo + --import-source on).
usage: python profiles/ncu_lines.py report.ncu-rep [top_n] > profiles/<name>_lines.md"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, hdr, lines = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; ci = {h: i for i, h in enumerate(hdr)}; continue
        if hdr is None or len(r) < len(hdr):
            continue
        def f(k):
            try: return float(r[ci[k]])
            except Exception: return 0.0
        if r[0].strip() != "":
            cur = (fname, int(r[0])); lines.setdefault(cur, dict(src=r[1].strip(), samples=0.0, inst=0.0, excess=0.0, stalls={}))
        d = lines[cur]
        # rows with an Address are SASS rows carrying the metrics
        if r[2].strip() == "":
            continue
        d["samples"] += f("# Samples"); d["inst"] += f("Instructions Executed"); d["excess"] += f("L1 Wavefronts Shared Excessive")
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                d["stalls"][k] = d["stalls"].get(k, 0.0) + f(k)
    tot = sum(d["samples"] for d in lines.values()) or 1.0
    toti = sum(d["inst"] for d in lines.values()) or 1.0
    print(f"# {path}: per-line warp-stall samples (total {tot:.0f}), warp instructions (total {toti:.0f})\n")
    print("| file:line | samples % | inst % | excess smem wavefronts | top stalls | source |\n|---|---|---|---|---|---|")
    for key, d in sorted(sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top], key=lambda kv: kv[0]):
        st = ", ".join(f"{k[6:]} {v / max(d['samples'], 1):.0%}" for k, v in sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3] if v > 0)
        print(f"| {key[0]}:{key[1]} | {100 * d['samples'] / tot:.1f} | {100 * d['inst'] / toti:.1f} | {d['excess']:.0f} | {st} | `{d['src'][:90]}` |")
    agg = {}
    for d in lines.values():
        for k, v in d["stalls"].items():
            agg[k] = agg.get(k, 0.0) + v
    print("\nstall totals: " + ", ".join(f"{k[6:]} {v / tot:.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
