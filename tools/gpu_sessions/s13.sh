set -x
mkdir -p gpurun_out/s13
CMD="python tools/prof_step.py --config humanoid --nenv 4096 --mode coop --launches 10 --warmup 100"
$CMD > gpurun_out/s13/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step_coop -s 3 -c 1 -o gpurun_out/s13/r2_coop_humanoid $CMD > gpurun_out/s13/ncu.log 2>&1
cat gpurun_out/s13/plain.log
