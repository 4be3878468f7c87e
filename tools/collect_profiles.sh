#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/collect_profiles.sh [tag]'): everything profiles/ is made from, into gpurun_out/.
# Afterwards, here:  ncu -i gpurun_out/prof_<tag>.ncu-rep --page raw --csv > gpurun_out/prof_<tag>_raw.csv
#                    python tools/ncu_summary.py gpurun_out/prof_<tag>_raw.csv profiles/<tag>_ncu_full_summary.json --traffic profiles/ncu_traffic.json --step
TAG=${1:-r2}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
# bench numbers first (never taken under a profiler), then the profiler passes of the same command
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo reference rc=$?
python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config2_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo rc=$?
python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config4_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo rc=$?
python bench.py --heads-mode fp32x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32x3_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a_$TAG.log 2>&1; echo rc=$?
if [ "$2" == "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:"heads_dw|heads_l1|heads_wide|nms_epoch" --launch-skip 20 -c 10 -f \
    -o gpurun_out/prof_$TAG python tools/time_config.py 384 1280 8 10 64 efficientdet-d0 fp16 > gpurun_out/ncu_b_$TAG.log 2>&1; echo rc=$?
# the stand-alone K2 kernels (fp64 reference arithmetic, fp32 per-tile, fp32 persistent TMA-staged) at configs[3], T = 10, B = 64
ncu --set full --clock-control none --import-source on -k regex:"decode_moments|decode_stream" --launch-skip 40 -c 6 -f \
    -o gpurun_out/prof_k2_$TAG python tools/postproc_sweep.py --Ts 10 --batches 64 --reps 1 > gpurun_out/ncu_k2_$TAG.log 2>&1; echo rc=$?
fi
# the other BASELINE configs
timeout 300 python tools/postproc_sweep.py --Ts 1,10,20,30 --out gpurun_out/postproc_sweep_$TAG.json > gpurun_out/postproc_sweep_$TAG.log 2>&1; echo rc=$?
timeout 200 python tools/autolabel_pass.py --images 512 --batch 16 --out gpurun_out/autolabel_pass_d2_$TAG.json > gpurun_out/autolabel_pass_$TAG.log 2>&1; echo rc=$?
python tools/time_config.py 512 512 8 10 1 efficientdet-d0 fp16 > gpurun_out/time_b1_$TAG.log 2>&1; tail -1 gpurun_out/time_b1_$TAG.log
python tools/time_nms.py > gpurun_out/time_nms_$TAG.log 2>&1; tail -3 gpurun_out/time_nms_$TAG.log
