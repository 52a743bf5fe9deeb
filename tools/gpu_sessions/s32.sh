set -x
mkdir -p gpurun_out/s32
CMD="python tools/prof_step.py --config cheetah --nenv 8192 --launches 10 --warmup 100"
$CMD > gpurun_out/s32/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step_spec -s 5 -c 1 -o gpurun_out/s32/r2_final_cheetah_spec $CMD > gpurun_out/s32/ncu_full.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s32/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s32/launches_bench_cheetah.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/s32/ncu_launches.log 2>&1
CMDH="python tools/prof_step.py --config humanoid --nenv 4096 --launches 10 --warmup 100"
$CMDH > gpurun_out/s32/plain_h.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_solve_coop -s 100 -c 1 -o gpurun_out/s32/r2_final_humanoid_solve $CMDH > gpurun_out/s32/ncu_full_h.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 30 --csv --log-file gpurun_out/s32/launches_humanoid.csv $CMDH > gpurun_out/s32/ncu_launches_h.log 2>&1
python bench.py --steps 300 --warmup 100 > gpurun_out/s32/bench_cheetah_full.json 2> gpurun_out/s32/bench_cheetah_full.err
python bench.py --config humanoid --steps 200 --warmup 100 > gpurun_out/s32/bench_humanoid_full.json 2> gpurun_out/s32/bench_humanoid_full.err
python bench.py --config humanoid --steps 200 --warmup 100 --ctrl-scale 0.125 --no-cpu-baseline > gpurun_out/s32/bench_humanoid_resting.json 2> /dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s32/bench_ref.json 2> gpurun_out/s32/bench_ref.err
tail -c 300 gpurun_out/s32/ncu_full.log; ls -la gpurun_out/s32
