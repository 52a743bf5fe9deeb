mkdir -p gpurun_out/s41
CMD2="python tools/prof_step.py --config cheetah --nenv 8192 --launches 10 --warmup 100"
$CMD2 > gpurun_out/s41/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step_spec -s 5 -c 1 -o gpurun_out/s41/r2_last_cheetah_spec $CMD2 > gpurun_out/s41/ncu.log 2>&1
tail -3 gpurun_out/s41/ncu.log
CMD3="python tools/prof_step.py --config humanoid --nenv 4096 --launches 6 --warmup 100"
$CMD3 > gpurun_out/s41/plain_h.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_solve_coop -s 3 -c 1 -o gpurun_out/s41/r2_last_humanoid_solve $CMD3 > gpurun_out/s41/ncu_h.log 2>&1
tail -3 gpurun_out/s41/ncu_h.log
