set -x
mkdir -p gpurun_out/s14
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/s14/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s14/pytest.log
tail -8 gpurun_out/s14/pytest.log
timeout 200 python bench.py --config humanoid --mode coop --steps 300 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s14/bench_humanoid_coop.json 2>/dev/null
timeout 200 python bench.py --config cheetah --mode coop --steps 300 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s14/bench_cheetah_coop.json 2>/dev/null
