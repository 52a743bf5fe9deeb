mkdir -p gpurun_out/s42
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_n3_n4.py tests/test_gpu_forward.py tests/test_abi.py tests/test_cpp_api.py tests/test_fluid.py -m gpu -q > gpurun_out/s42/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s42/pytest.log
tail -4 gpurun_out/s42/pytest.log
