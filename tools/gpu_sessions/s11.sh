set -x
mkdir -p gpurun_out/s11
timeout 300 python bench.py --steps 500 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s11/bench_cheetah.json 2> gpurun_out/s11/bench_cheetah.err
timeout 300 python bench.py --config humanoid --steps 300 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s11/bench_humanoid.json 2> gpurun_out/s11/bench_humanoid.err
python - <<'PY'
import json
for f in ("gpurun_out/s11/bench_cheetah.json","gpurun_out/s11/bench_humanoid.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, d["value"], d["ms_per_step"], d["value_resident_one_launch"])
PY
