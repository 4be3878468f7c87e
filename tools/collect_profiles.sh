#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/collect_profiles.sh'): everything profiles/ is made from, into gpurun_out/.
# Afterwards, here:  ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv
#                    python tools/ncu_summary.py gpurun_out/prof_final_raw.csv profiles/rN_final_ncu_full_summary.json --traffic profiles/ncu_traffic.json
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
# bench numbers first (never taken under a profiler), then the profiler passes of the same command
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"heads_fused|heads_ig|heads_l1|heads_wide|nms_v5_sorted" -c 10 -f \
    -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; echo rc=$?
# the other BASELINE configs
timeout 300 python tools/postproc_sweep.py --Ts 1,10,20,30 --out gpurun_out/postproc_sweep.json > gpurun_out/postproc_sweep.log 2>&1; echo rc=$?
timeout 200 python tools/autolabel_pass.py --images 512 --batch 16 --out gpurun_out/autolabel_pass_d2.json > gpurun_out/autolabel_pass.log 2>&1; echo rc=$?
python tools/time_config.py 720 1280 10 20 32 > gpurun_out/time_bdd.log 2>&1; tail -1 gpurun_out/time_bdd.log
python tools/time_config.py 768 768 10 30 16 efficientdet-d2 bf16 > gpurun_out/time_d2.log 2>&1; tail -1 gpurun_out/time_d2.log
python tools/time_config.py 512 512 8 10 1 > gpurun_out/time_b1.log 2>&1; tail -1 gpurun_out/time_b1.log
