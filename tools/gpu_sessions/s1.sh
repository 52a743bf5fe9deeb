set -x
mkdir -p gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/s1/smi.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1/pytest.log
timeout 300 python bench.py --steps 1000 --warmup 100 > gpurun_out/s1/bench_cheetah.json 2> gpurun_out/s1/bench_cheetah.err
for L in 16 8 4; do
 timeout 120 python bench.py --steps 500 --warmup 100 --lanes $L --no-e2e --no-cpu-baseline > gpurun_out/s1/bench_cheetah_l$L.json 2>&1
 timeout 120 python bench.py --steps 500 --warmup 100 --lanes $L --no-e2e --no-cpu-baseline --no-flush > gpurun_out/s1/bench_cheetah_l${L}_nf.json 2>&1
done
timeout 120 python bench.py --steps 500 --warmup 100 --no-e2e --no-cpu-baseline --no-flush > gpurun_out/s1/bench_cheetah_l32_nf.json 2>&1
for B in 64 128; do
 timeout 120 python bench.py --steps 500 --warmup 100 --lanes 8 --block $B --no-e2e --no-cpu-baseline --no-flush > gpurun_out/s1/bench_cheetah_l8_b${B}_nf.json 2>&1
done
timeout 300 python bench.py --config humanoid --steps 300 --warmup 100 > gpurun_out/s1/bench_humanoid.json 2> gpurun_out/s1/bench_humanoid.err
tail -3 gpurun_out/s1/pytest.log
