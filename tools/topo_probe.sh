#!/usr/bin/env bash
# What the GPU box exposes about CPU / NUMA / PCIe topology (for the e2e figure's host side).
echo "== cpus: $(nproc) ; nodes: $(ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l)"
for n in /sys/devices/system/node/node*; do echo "$(basename $n): cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4,$5}')"; done 2>/dev/null
nvidia-smi topo -m 2>&1 | head -30
python - <<'PY'
import os, pynvml
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    w = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    print("gpu", i, "cpu affinity words", [hex(int(x)) for x in w], "pcie gen/width", pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
print("allowed", sorted(os.sched_getaffinity(0)))
PY
