set -x
mkdir -p gpurun_out/s19
CMD2="python tools/prof_step.py --config cheetah --nenv 8192 --launches 10 --warmup 100"
$CMD2 > gpurun_out/s19/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step_spec -s 5 -c 1 -o gpurun_out/s19/r2b_cheetah_spec $CMD2 > gpurun_out/s19/ncu2.log 2>&1
tail -3 gpurun_out/s19/ncu2.log
