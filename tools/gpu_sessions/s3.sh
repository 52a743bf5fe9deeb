set -x
mkdir -p gpurun_out/s3
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/s3/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3/pytest.log
tail -40 gpurun_out/s3/pytest.log
timeout 300 python bench.py --steps 500 --warmup 100 --no-cpu-baseline > gpurun_out/s3/bench_cheetah.json 2> gpurun_out/s3/bench_cheetah.err
