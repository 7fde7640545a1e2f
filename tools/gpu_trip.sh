#!/bin/bash
# GPU-box driver used through gpurun: unit tests, isolated tensor-core probe, whole-model parity. Logs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/nvidia_smi.txt
STAGES="${1:-kernels probe fp32 bf16}"
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
for s in $STAGES; do
  case $s in
    kernels) timeout 900 $PT tests/test_kernels_gpu.py -x > gpurun_out/kernels.log 2>&1 ;;
    kernels_all) timeout 900 $PT tests/test_kernels_gpu.py > gpurun_out/kernels.log 2>&1 ;;
    probe) timeout 1200 python tools/tc_probe.py > gpurun_out/tc_probe.log 2>&1 ;;
    fp32) CUDA_LAUNCH_BLOCKING=1 timeout 1200 $PT tests/test_model_gpu.py -k fp32 > gpurun_out/model_fp32.log 2>&1 ;;
    bf16) timeout 1200 $PT tests/test_model_gpu.py -k bf16 -s > gpurun_out/model_bf16.log 2>&1 ;;
    tc) timeout 1200 $PT tests/test_tc_gpu.py > gpurun_out/tc.log 2>&1 ;;
    all) timeout 2400 $PT tests > gpurun_out/all.log 2>&1 ;;
  esac
done
for f in gpurun_out/*.log; do echo "=== $f"; tail -15 $f; done
