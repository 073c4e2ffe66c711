# ncu launch lists of the final round-2 build (gpurun -- 'bash tools/r2_ncu_lists.sh'); summarised with tools/summarize_launches.py
set -x
export INFUR_B200_NO_GRAPH=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_f16_b8_1080p.csv python tools/profile_step.py --iters 0 --steps 1 --cuda-profiler > gpurun_out/r2_ncu_step_f16.log 2>&1
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_int8_b8_1080p.csv python tools/profile_step.py --kind fcn50_int8 --iters 0 --steps 1 --cuda-profiler > gpurun_out/r2_ncu_step_int8.log 2>&1
INFUR_BENCH_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --model f16 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out/r2_launches_*.csv
