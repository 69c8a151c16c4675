#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.md   (needs `ncu` on PATH; no GPU)"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "CTAs/SM allowed by registers"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM allowed by smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA-pipe instructions % of peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "TENSOR pipe cycles active % (any tensor sub-pipe)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor sub-pipe HMMA cycles active %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor-pipe instructions % of peak"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions (UTCHMMA / UBLKCP issue)"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tensor-memory pipe (LDTM / STTM) % of peak"),
    ("lts__t_bytes.sum", "L2 bytes (all traffic)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "shared-memory pipe % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("idc__request_hit_rate.pct", "constant cache hit rate %"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("sm__sass_inst_executed_op_local_ld.sum", "local (spill) loads"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}\n")
    for k, r in enumerate(data):
        print(f"## launch {k}: {r[col['Kernel Name']][:110]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            if key in col:
                print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        extra = [h for h in hdr if ("tensor" in h or "tmem" in h) and h not in dict(KEYS) and ".avg.pct_of_peak_sustained_active" in h]
        for h in extra:
            print(f"| `{h}` | {r[col[h]]} | {units[col[h]]} |")
        print("\nwarp stall reasons (cycles stalled per issued instruction):\n")
        st = [(h.split("issue_stalled_")[1].split("_per_issue")[0], float(r[i])) for h, i in col.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        for name, v in sorted(st, key=lambda t: -t[1])[:8]:
            print(f"* {name}: {v:.3f}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
