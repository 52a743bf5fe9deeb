mkdir -p gpurun_out/s16
run() { for L in 0 1; do for cfg in humanoid cheetah; do
  OX_B200_COOP_LOCKSTEP=$L timeout 200 python bench.py --config $cfg --mode coop --steps 300 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s16/bench_${cfg}_coop_$1_l$L.json 2>/dev/null
done; done; }
run t128
OX_B200_COOP_LOCKSTEP=1 timeout 300 python -m pytest tests/test_gpu_coop.py -q 2>&1 | tail -2
cp oxide_control_b200/lib/libox_b200_c256.so oxide_control_b200/lib/libox_b200.so
run t256
python - <<'PY'
import json
for t in ("t128","t256"):
 for L in (0,1):
  for c in ("humanoid","cheetah"):
    try:
      d=json.loads([l for l in open(f"gpurun_out/s16/bench_{c}_coop_{t}_l{L}.json") if l.startswith("{")][-1]); print(t,"lockstep",L,c, "%.3f ms"%d["ms_per_step"], "%.2fM"%(d["value"]/1e6), "resident %.2fM"%(d["value_resident_one_launch"]/1e6))
    except Exception as e: print(t,L,c,"ERR",e)
PY
