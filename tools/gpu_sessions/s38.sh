mkdir -p gpurun_out/s38
timeout 1500 python -m pytest tests/test_zoo_parity.py tests/test_golden.py tests/test_touch.py tests/test_force_torque.py tests/test_jit.py -m gpu -q -k "zoo_a or jit or touch or force" > gpurun_out/s38/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s38/pytest.log
tail -6 gpurun_out/s38/pytest.log
