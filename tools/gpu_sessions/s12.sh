set -x
mkdir -p gpurun_out/s12
timeout 600 python -m pytest tests/test_gpu_coop.py -x -q > gpurun_out/s12/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s12/pytest.log
tail -40 gpurun_out/s12/pytest.log
for cfg in humanoid cheetah; do
  timeout 200 python bench.py --config $cfg --mode coop --steps 300 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s12/bench_${cfg}_coop.json 2> gpurun_out/s12/bench_${cfg}_coop.err
done
python - <<'PY'
import json
for c in ("humanoid","cheetah"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/s12/bench_{c}_coop.json") if l.startswith("{")][-1]); print(c, d["value"], d["ms_per_step"], d["value_resident_one_launch"], d["mean_solver_iters"], d["mean_ncon"])
    except Exception as e: print(c, "ERR", e, open(f"gpurun_out/s12/bench_{c}_coop.err").read()[-600:])
PY
