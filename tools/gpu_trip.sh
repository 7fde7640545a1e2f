#!/bin/bash
# GPU-box driver used through gpurun: unit tests, isolated tensor-core probe, whole-model parity. Logs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/nvidia_smi.txt
STAGES="${1:-kernels probe fp32 bf16}"
for s in $STAGES; do
  case $s in
    kernels) timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/kernels.log ;;
    kernels_all) timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/kernels.log ;;
    probe) timeout 1200 python tools/tc_probe.py > gpurun_out/tc_probe.log 2>&1 ;;
    fp32) timeout 1200 python -m pytest tests/test_model_gpu.py -q -m gpu -k fp32 --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/model_fp32.log ;;
    bf16) timeout 1200 python -m pytest tests/test_model_gpu.py -q -m gpu -k bf16 -s --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/model_bf16.log ;;
    tc) timeout 1200 python -m pytest tests/test_tc_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/tc.log ;;
    all) timeout 2400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/all.log ;;
  esac
done
for f in gpurun_out/*.log; do echo "=== $f"; tail -25 $f; done
