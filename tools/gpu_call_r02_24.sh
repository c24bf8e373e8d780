#!/usr/bin/env bash
# ncu on the library yardsticks at the config-2 shape (non-causal and causal): how do FA4 / cuDNN spend the SM?
cd "$(dirname "$0")/.."
out=gpurun_out/r02_24; mkdir -p $out
for prov in fa4 cudnn ours; do
for causal in 0 1; do
cat > /tmp/one.py <<PY
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import yardstick
f = yardstick.providers()["$prov"] if hasattr(yardstick, "providers") else None
PY
python - <<PY > $out/run_${prov}_${causal}.py
print('''
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "flash-attention-v100_b200"))
prov, causal = "$prov", bool($causal)
if prov == "cudnn":
    from torch.nn.attention import SDPBackend, sdpa_kernel
    import torch.nn.functional as F
    def fn(q, k, v):
        with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
            return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal).transpose(1, 2)
elif prov == "fa4":
    from vllm.vllm_flash_attn.cute.interface import flash_attn_func as f4
    def fn(q, k, v):
        return f4(q, k, v, causal=causal)
else:
    from flash_attn_v100 import flash_attn_func
    def fn(q, k, v):
        return flash_attn_func(q, k, v, causal=causal)
q, k, v = (torch.randn(8, 4096, 32, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(4): fn(q, k, v)
torch.cuda.synchronize()
''')
PY
timeout 300 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section LaunchStats --section Occupancy --metrics smsp__inst_executed.sum,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__shared_mem_per_block_dynamic --clock-control none -k "regex:(fmha|attention|fa_fwd|flash|kernel_cutlass|cudnn)" -s 3 -c 1 --csv --page raw --log-file $out/ncu_${prov}_${causal}.csv python $out/run_${prov}_${causal}.py > $out/log_${prov}_${causal}.txt 2>&1
done; done
python - <<'PY'
import csv, glob
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
for f in sorted(glob.glob("gpurun_out/r02_24/ncu_*.csv")):
    rows = [r for r in csv.reader(open(f)) if r]
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r]
    if not h:
        print(f, "no kernel captured"); continue
    hdr = rows[h[0]]
    for r in rows[h[0] + 2:]:
        d = dict(zip(hdr, r))
        print(f.split("/")[-1], " | ".join(f"{k.split('.')[0][-28:]}={d.get(k, '?')[:48]}" for k in keys))
PY
