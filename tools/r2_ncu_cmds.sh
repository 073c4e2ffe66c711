# Round-2 profiling pass (run on a GPU box: gpurun -- 'bash tools/r2_ncu_cmds.sh'); results land in gpurun_out/ and are summarised
# into profiles/ with tools/summarize_launches.py and tools/summarize_ncu.py.  ncu runs use plain launches (INFUR_B200_NO_GRAPH=1).
set -x
export INFUR_B200_NO_GRAPH=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_f16_b8_1080p.csv python tools/profile_step.py --iters 0 --steps 1 --cuda-profiler > gpurun_out/r2_ncu_step_f16.log 2>&1
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_int8_b8_1080p.csv python tools/profile_step.py --kind fcn50_int8 --iters 0 --steps 1 --cuda-profiler > gpurun_out/r2_ncu_step_int8.log 2>&1
INFUR_BENCH_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --model f16 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'post_cell_kernel|pre_unit_warp128|frame_blend' -c 3 -o gpurun_out/r2_ncu_prepost -f python tools/profile_step.py --iters 0 --steps 1 --cuda-profiler > gpurun_out/r2_ncu_prepost.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_b2b_kernel|stem_pool_kernel' -c 3 -o gpurun_out/r2_ncu_fused -f python tools/profile_step.py --iters 0 --steps 1 --cuda-profiler > gpurun_out/r2_ncu_fused.log 2>&1
unset INFUR_B200_NO_GRAPH
python tools/profile_step.py --kind fcn50_int8 --iters 5 > gpurun_out/r2_events_int8.txt 2>&1
python tools/profile_step.py --iters 5 > gpurun_out/r2_events_f16.txt 2>&1
tail -n 3 gpurun_out/r2_events_f16.txt; tail -n 3 gpurun_out/r2_events_int8.txt
