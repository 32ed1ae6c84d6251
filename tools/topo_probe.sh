#!/bin/bash
# host topology of the GPU box, for placing page-locked frame buffers (bench.py: bind_to_gpu_numa_node)
nvidia-smi topo -m 2>&1 | head -30
echo "--- numa nodes"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist 2>/dev/null) mem $(grep MemTotal $n/meminfo 2>/dev/null | awk '{print $4}') kB"; done
echo "--- gpu numa_node (sysfs)"
for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do bb=$(echo $b | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "$b $(cat /sys/bus/pci/devices/$bb/numa_node 2>&1)"; done
echo "--- this process"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
nproc; which numactl
python - <<'PY'
import pynvml
pynvml.nvmlInit()
for i in range(pynvml.nvmlDeviceGetCount()):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    try:
        print(i, "affinity", [hex(x) for x in pynvml.nvmlDeviceGetCpuAffinity(h, 8)])
    except Exception as e:
        print(i, "affinity err", e)
    try:
        print(i, "numa", pynvml.nvmlDeviceGetNumaNodeId(h))
    except Exception as e:
        print(i, "numa err", e)
PY
