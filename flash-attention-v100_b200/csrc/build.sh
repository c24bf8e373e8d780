#!/usr/bin/env bash
# Builds libfa_b200.so (the C-ABI library of include/fa_b200.h) for sm_100a, in-tree.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../lib"
mkdir -p "${out}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
python3 "${here}/gen_tmem_ldst.py" "${here}/tmem_ldst_gen.cuh"
python3 "${here}/gen_umma_issue.py" "${here}/umma_issue_gen.cuh"
# FA_VARIANT_FLAGS / FA_LIB_NAME: build an A/B variant (e.g. -DFA_EMU_COUNT=0) next to the default library
LIB_NAME="${FA_LIB_NAME:-libfa_b200.so}"
"${NVCC}" -std=c++17 -O3 -lineinfo ${FA_VARIANT_FLAGS:-} \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr \
  -Xptxas -v \
  -shared -cudart static \
  -o "${out}/${LIB_NAME}" "${here}/fa_b200_api.cu" 2> "${out}/ptxas.log" || { cat "${out}/ptxas.log"; exit 1; }
grep -E "error|warning" "${out}/ptxas.log" | grep -v "Wno" | head -20 || true
echo "built ${out}/${LIB_NAME}"
