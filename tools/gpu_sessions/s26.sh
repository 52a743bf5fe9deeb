mkdir -p gpurun_out/s26
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fp32.py -m gpu -q -x 2>&1 | tail -4
for sc in 1 0.125; do
timeout 300 python bench.py --config humanoid --steps 300 --warmup 100 --no-cpu-baseline --no-e2e --ctrl-scale $sc > gpurun_out/s26/hum_s$sc.json 2>gpurun_out/s26/hum_s$sc.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s26/hum_s$sc.json") if l.startswith("{")][-1])
print("humanoid scale $sc", "%.4f ms"%d["ms_per_step"], "%.2fM"%(d["value"]/1e6), "resident %.2fM"%(d.get("value_resident_one_launch",0)/1e6))
PY
done
