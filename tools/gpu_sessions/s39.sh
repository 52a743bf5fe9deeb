mkdir -p gpurun_out/s39
# final record of the round on the final build: driver-protocol bench (defaults), the reference arm, humanoid, full GPU suite, smoke
timeout 600 python bench.py > gpurun_out/s39/bench_cheetah_full.json 2> gpurun_out/s39/bench_cheetah_full.err; tail -1 gpurun_out/s39/bench_cheetah_full.json | cut -c1-240
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s39/bench_cheetah_driver_protocol.json 2> /dev/null; tail -1 gpurun_out/s39/bench_cheetah_driver_protocol.json | cut -c1-240
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/s39/bench_ref.json 2> /dev/null; tail -1 gpurun_out/s39/bench_ref.json | cut -c1-240
timeout 600 python bench.py --config humanoid --steps 300 --warmup 100 > gpurun_out/s39/bench_humanoid_full.json 2> /dev/null; tail -1 gpurun_out/s39/bench_humanoid_full.json | cut -c1-240
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/s39/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s39/pytest_gpu.log
tail -6 gpurun_out/s39/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s39/launches_bench_cheetah.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s39/ncu.log 2>&1; echo "ncu rc=$?"
