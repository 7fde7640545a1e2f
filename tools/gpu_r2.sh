#!/bin/bash
# Round-2 GPU-box trips.  Usage (through gpurun): bash tools/gpu_r2.sh [stages...]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/nvidia_smi.txt
STAGES="${*:-tests smoke bench}"
NCU=/usr/local/cuda/bin/ncu
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for s in $STAGES; do
  echo "=== stage $s $(date +%T)"
  case $s in
    tests) timeout 2400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --tb=short --maxfail=40 -s > gpurun_out/tests.log 2>&1; tail -60 gpurun_out/tests.log | cut -c1-400 ;;
    tests_x) timeout 2400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --tb=short -x > gpurun_out/tests.log 2>&1; tail -30 gpurun_out/tests.log | cut -c1-400 ;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log ;;
    bench) timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 2500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err ;;
    bench_quick) timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python tools/show_bench.py gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err ;;
    bench_pdl) XV2_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; python tools/show_bench.py gpurun_out/bench_pdl.json | head -8; tail -5 gpurun_out/bench_pdl.err ;;
    bench_c3|bench_c4|bench_c5) c=${s#bench_}; timeout 1500 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; python tools/show_bench.py gpurun_out/bench_$c.json; tail -5 gpurun_out/bench_$c.err ;;
    ref) timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.json ;;
    layers) timeout 600 python tools/layer_profile.py > gpurun_out/layers.txt 2> gpurun_out/layers.err; head -3 gpurun_out/layers.txt; tail -3 gpurun_out/layers.err ;;
    launches) timeout 1200 $NCU --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/layer_profile.py --ncu > gpurun_out/launches.log 2>&1; wc -l gpurun_out/launches.csv ;;
    traffic) timeout 1500 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/traffic.csv python tools/layer_profile.py --ncu > gpurun_out/traffic.log 2>&1; wc -l gpurun_out/traffic.csv ;;
    ncu_zoo) # one --set full row per kernel family at its C2 shape; the report stays on the box, the CSV comes back
        timeout 1200 $NCU --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:"xv2::" -c ${ZOO_COUNT:-120} \
          -f -o /tmp/prof_zoo python tools/kernel_zoo.py ${ZOO_ARGS:-} > gpurun_out/ncu_zoo${ZOO_TAG:-}.log 2>&1; tail -3 gpurun_out/ncu_zoo${ZOO_TAG:-}.log
        $NCU -i /tmp/prof_zoo.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_bytes.sum > gpurun_out/ncu_zoo${ZOO_TAG:-}.csv 2>/dev/null; wc -l gpurun_out/ncu_zoo${ZOO_TAG:-}.csv
        $NCU -i /tmp/prof_zoo.ncu-rep --page details --csv --section WarpStateStats --section SpeedOfLight > gpurun_out/ncu_zoo${ZOO_TAG:-}_details.csv 2>/dev/null; wc -c gpurun_out/ncu_zoo${ZOO_TAG:-}_details.csv ;;
    scale2|scale4|scale8) n=${s#scale}; for c in ${CFGS:-c2}; do
        timeout -k 10 ${SCALE_TIMEOUT:-240} $TR --nproc-per-node $n --master-port 295$n bench.py --config $c --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_${c}_n$n${TAG:-}.json 2> gpurun_out/scale_${c}_n$n${TAG:-}.err
        python tools/show_bench.py gpurun_out/scale_${c}_n$n${TAG:-}.json | head -3; tail -3 gpurun_out/scale_${c}_n$n${TAG:-}.err | cut -c1-300; done ;;
    layers_wg) for v in "64 128" "128 128" "64 256"; do set -- $v; XV2_WG_PIX=$1 XV2_WG_BN=$2 timeout 600 python tools/layer_profile.py > gpurun_out/layers_wg_$1_$2.txt 2> gpurun_out/layers.err
        echo "--- XV2_WG_PIX=$1 XV2_WG_BN=$2"; python tools/layer_sum.py gpurun_out/layers_wg_$1_$2.txt -n 12; tail -2 gpurun_out/layers.err; done ;;
    dbg_fork) timeout 300 python tools/dbg_fork.py > gpurun_out/dbg_fork.txt 2>&1; tail -12 gpurun_out/dbg_fork.txt ;;
    tests_sel) timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_blocks_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -q --no-header -p no:cacheprovider --tb=short -k "${TESTS_K:-pools or convt or fork or bottleneck or batch_norm or model}" > gpurun_out/tests_sel.log 2>&1; tail -25 gpurun_out/tests_sel.log | cut -c1-300 ;;
    loader) timeout 600 python tools/loader_bench.py --tiles 48 > gpurun_out/loader_bench.txt 2> gpurun_out/loader_bench.err; tail -2 gpurun_out/loader_bench.txt; tail -3 gpurun_out/loader_bench.err ;;
    tests_new) timeout 900 python -m pytest tests/test_augment_gpu.py tests/test_postprocess_gpu.py -q --no-header -p no:cacheprovider --tb=short > gpurun_out/tests_new.log 2>&1; tail -30 gpurun_out/tests_new.log | cut -c1-300 ;;
    *) echo "unknown stage $s" ;;
  esac
done
echo "=== done $(date +%T)"
