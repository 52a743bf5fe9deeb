set -x
mkdir -p gpurun_out/s7
CMD="python tools/prof_step.py --config humanoid --nenv 4096 --launches 10 --warmup 100"
$CMD > gpurun_out/s7/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 30 --csv --log-file gpurun_out/s7/launches_humanoid.csv $CMD > gpurun_out/s7/ncu1.log 2>&1
$CMD > gpurun_out/s7/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step_spec -s 200 -c 2 -o gpurun_out/s7/r2_humanoid_prepost $CMD > gpurun_out/s7/ncu2.log 2>&1
cat gpurun_out/s7/plain.log; tail -3 gpurun_out/s7/ncu1.log gpurun_out/s7/ncu2.log
