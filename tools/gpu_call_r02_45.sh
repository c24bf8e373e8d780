#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_45
bash tools/ncu_capture.sh gpurun_out/r02_45/ncu_fwd_d256 fa_fwd_sm100_kernel 2 python tools/profile_target.py d256 4
ls -la gpurun_out/r02_45
