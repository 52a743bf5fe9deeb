set -x
mkdir -p gpurun_out/s6
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/s6/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6/pytest.log
tail -5 gpurun_out/s6/pytest.log
timeout 300 python bench.py --config humanoid --steps 300 --warmup 100 --no-cpu-baseline > gpurun_out/s6/bench_humanoid.json 2> gpurun_out/s6/bench_humanoid.err
timeout 300 python bench.py --config humanoid --steps 300 --warmup 300 --ctrl-scale 0.125 --no-cpu-baseline > gpurun_out/s6/bench_humanoid_rest.json 2> gpurun_out/s6/bench_humanoid_rest.err
timeout 300 python bench.py --steps 500 --warmup 100 --no-cpu-baseline > gpurun_out/s6/bench_cheetah.json 2> gpurun_out/s6/bench_cheetah.err
timeout 200 python tools/sanitize_run.py > gpurun_out/s6/sanitize_plain.log 2>&1; echo "rc=$?" >> gpurun_out/s6/sanitize_plain.log; tail -3 gpurun_out/s6/sanitize_plain.log
