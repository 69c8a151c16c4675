#!/bin/bash
# One GPU-box pass of the round's evidence: GPU tests, the bench line (+ the reference arm), the ncu launch list of the bench command and
# full ncu captures of the dominant kernels, SUMMARISED ON THE BOX (the .ncu-rep files exceed what gpurun copies back).
#   gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh r2j'
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/${TAG}_tests.log
tail -3 $O/${TAG}_tests.log
python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -2 $O/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
cap() {   # name, kernel regex, skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/${name} "$@" > $O/${TAG}_ncu_${name}.log 2>&1
    if [ -f /tmp/${name}.ncu-rep ]; then
        python profiles/summarize_ncu.py /tmp/${name}.ncu-rep > $O/${TAG}_${name}_ncu.md 2>> $O/${TAG}_ncu_${name}.log
        python profiles/ncu_lines.py /tmp/${name}.ncu-rep 45 > $O/${TAG}_${name}_lines.md 2>> $O/${TAG}_ncu_${name}.log
    fi
}
cap wide_kernel nempc_wide_kernel 9 python tools/wide_check.py --no-check --time 2048
cap fast64_kernel nempc_fast64_kernel 2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-solver
cap fast_kernel nempc_fast_kernel 4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-solver --no-side-workloads
cap tc_kernel nempc_tc_kernel 1 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline
cap dmma_net_kernel nempc_dmma_net 5 python tools/time_eval.py --workload C3 --compute float64 --batch 2048 --reps 1
ls -la $O | tail -20
