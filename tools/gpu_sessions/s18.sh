mkdir -p gpurun_out/s18
python -m pytest tests/test_equality_mocap.py -m gpu -q -s 2>&1 | grep "f32\|passed\|failed\|Error" | tail -30
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s18/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s18/pytest_gpu.log
tail -8 gpurun_out/s18/pytest_gpu.log
