mkdir -p gpurun_out/s36
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/s36/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s36/pytest_gpu.log
tail -8 gpurun_out/s36/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
