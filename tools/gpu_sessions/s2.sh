set -x
mkdir -p gpurun_out/s2
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/s2/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2/pytest.log
tail -30 gpurun_out/s2/pytest.log
