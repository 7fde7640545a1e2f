#!/bin/bash
# One GPU-box trip: parity tests, bench line, per-launch table, ncu launch list + one full capture of the top kernels.
# Usage (through gpurun): bash tools/gpu_round.sh [stages...]   stages: tests bench layers launches ncu_conv ncu_wgrad
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/nvidia_smi.txt
STAGES="${*:-tests bench layers launches ncu_conv}"
NCU=/usr/local/cuda/bin/ncu
for s in $STAGES; do
  echo "=== stage $s $(date +%T)"
  case $s in
    tests) timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --tb=short -x > gpurun_out/tests.log 2>&1; tail -5 gpurun_out/tests.log ;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log ;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
    layers) timeout 600 python tools/layer_profile.py > gpurun_out/layers.txt 2> gpurun_out/layers.err; head -3 gpurun_out/layers.txt; tail -3 gpurun_out/layers.err ;;
    launches) timeout 1200 $NCU --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/layer_profile.py --ncu > gpurun_out/launches.log 2>&1; wc -l gpurun_out/launches.csv ;;
    launches_bench) timeout 1500 $NCU --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
        --log-file gpurun_out/launches_bench.csv env XV2_NO_GRAPH=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; wc -l gpurun_out/launches_bench.csv ;;
    ncu_all) timeout 1500 $NCU --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:"conv_strip_kernel|conv_tc_kernel|wgrad_strip_kernel|wgrad_tc_kernel|bn_stream_kernel" -s 180 -c 14 -f -o gpurun_out/prof_all python tools/layer_profile.py --ncu > gpurun_out/ncu_all.log 2>&1; tail -2 gpurun_out/ncu_all.log ;;
    ncu_conv) timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:conv_tc_kernel -s 0 -c 4 -f -o gpurun_out/prof_conv python tools/layer_profile.py --ncu > gpurun_out/ncu_conv.log 2>&1; tail -3 gpurun_out/ncu_conv.log ;;
    ncu_strip) timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:conv_strip_kernel -c 8 -f -o gpurun_out/prof_strip python tools/layer_profile.py --ncu > gpurun_out/ncu_strip.log 2>&1; tail -3 gpurun_out/ncu_strip.log ;;
    ncu_bn) timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:"bn_bwd" -c 8 -f -o gpurun_out/prof_bnbwd python tools/layer_profile.py --ncu > gpurun_out/ncu_bn.log 2>&1; tail -2 gpurun_out/ncu_bn.log
        timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:"bn_apply|bn_stats" -c 8 -f -o gpurun_out/prof_bnfwd python tools/layer_profile.py --ncu > gpurun_out/ncu_bn2.log 2>&1; tail -2 gpurun_out/ncu_bn2.log ;;
    traffic) timeout 1500 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/traffic.csv python tools/layer_profile.py --ncu > gpurun_out/traffic.log 2>&1; wc -l gpurun_out/traffic.csv ;;
    bench2) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -c 600 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err ;;
    ncu_wgrad) timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:wgrad_ -s 4 -c 6 -f -o gpurun_out/prof_wgrad python tools/layer_profile.py --ncu > gpurun_out/ncu_wgrad.log 2>&1; tail -3 gpurun_out/ncu_wgrad.log ;;
  esac
done
echo "=== done $(date +%T)"
