nvidia-smi topo -m 2>&1 | head -14
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
