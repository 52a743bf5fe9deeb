mkdir -p gpurun_out/s15
for L in 0 1; do for cfg in humanoid cheetah; do
  OX_B200_COOP_LOCKSTEP=$L timeout 200 python bench.py --config $cfg --mode coop --steps 300 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s15/bench_${cfg}_coop_l$L.json 2>/dev/null
done; done
OX_B200_COOP_LOCKSTEP=1 timeout 300 python -m pytest tests/test_gpu_coop.py -q 2>&1 | tail -2
python - <<'PY'
import json
for L in (0,1):
  for c in ("humanoid","cheetah"):
    d=json.loads([l for l in open(f"gpurun_out/s15/bench_{c}_coop_l{L}.json") if l.startswith("{")][-1]); print("lockstep",L,c, "%.3f ms"%d["ms_per_step"], "%.2fM"%(d["value"]/1e6), "resident %.2fM"%(d["value_resident_one_launch"]/1e6))
PY
