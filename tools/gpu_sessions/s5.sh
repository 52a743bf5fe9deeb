set -x
mkdir -p gpurun_out/s5
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/s5/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5/pytest.log
tail -30 gpurun_out/s5/pytest.log
timeout 300 python tools/sanitize_run.py > gpurun_out/s5/sanitize_plain.log 2>&1; echo "rc=$?" >> gpurun_out/s5/sanitize_plain.log; tail -5 gpurun_out/s5/sanitize_plain.log
timeout 300 python bench.py --steps 1000 --warmup 100 > gpurun_out/s5/bench_cheetah.json 2> gpurun_out/s5/bench_cheetah.err
timeout 300 python bench.py --config humanoid --steps 300 --warmup 100 > gpurun_out/s5/bench_humanoid.json 2> gpurun_out/s5/bench_humanoid.err
timeout 300 python bench.py --config humanoid --steps 300 --warmup 300 --ctrl-scale 0.125 > gpurun_out/s5/bench_humanoid_rest.json 2> gpurun_out/s5/bench_humanoid_rest.err
timeout 120 python tools/group_bench.py --gpus 1 --steps 500 > gpurun_out/s5/group_n1.json 2>&1
tail -2 gpurun_out/s5/*.err
