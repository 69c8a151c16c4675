# solver time of the bench workload against the shared-memory budget of the staged KKT kernel (NEMPC_KKT_SMEM_KB, default 192)
for kb in 128 160 192 208 224; do
  NEMPC_KKT_SMEM_KB=$kb python bench.py --no-cpu-baseline --no-side-workloads --steps 50 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($kb, d['mpc_solves']['value'], d['mpc_solves']['ms_per_batch'])"
done
