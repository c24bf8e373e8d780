#!/usr/bin/env bash
# Builds libfa_b200.so (the C-ABI library of include/fa_b200.h) for sm_100a, in-tree.
#   FA_VARIANT_FLAGS / FA_LIB_NAME : build an A/B variant (e.g. -DFA_EMU_COUNT=0) next to the default library
#   FA_BUILD_JITTER=1              : also build libfa_b200_jitter.so (-DFA_JITTER: random busy-waits before every mbarrier
#                                    wait / arrival; tests/test_gpu_protocol_stress.py runs parity cases against it), in parallel
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../lib"
mkdir -p "${out}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
python3 "${here}/gen_tmem_ldst.py" "${here}/tmem_ldst_gen.cuh"
python3 "${here}/gen_umma_issue.py" "${here}/umma_issue_gen.cuh"
build_one() {  # <extra flags> <library name> <log file>
  "${NVCC}" -std=c++17 -O3 -lineinfo $1 \
    -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr \
    -Xptxas -v \
    -shared -cudart static \
    -o "${out}/$2" "${here}/fa_b200_api.cu" 2> "$3" || { cat "$3"; return 1; }
  grep -E "error|warning" "$3" | grep -v "Wno" | head -20 || true
  echo "built ${out}/$2"
}
LIB_NAME="${FA_LIB_NAME:-libfa_b200.so}"
pids=()
if [ "${FA_BUILD_JITTER:-0}" = "1" ]; then
  build_one "-DFA_JITTER" "libfa_b200_jitter.so" "${out}/ptxas_jitter.log" &
  pids+=($!)
fi
build_one "${FA_VARIANT_FLAGS:-}" "${LIB_NAME}" "${out}/ptxas.log"
for pid in "${pids[@]:-}"; do [ -n "${pid}" ] && wait "${pid}"; done
