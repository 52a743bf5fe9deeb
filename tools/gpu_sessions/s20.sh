mkdir -p gpurun_out/s20
for L in 32 16 8 4 2; do
  for sc in 1 0.125; do
  timeout 300 python bench.py --config humanoid --steps 200 --warmup 100 --no-cpu-baseline --no-e2e --lanes $L --ctrl-scale $sc > gpurun_out/s20/hum_l${L}_s${sc}.json 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s20/hum_l${L}_s${sc}.json") if l.startswith("{")][-1])
print("lanes $L scale $sc", "%.3f ms"%d["ms_per_step"], "%.2fM"%(d["value"]/1e6), {k:d.get(k) for k in ("mean_ncon","mean_nefc","mean_niter") if k in d})
PY
  done
done
