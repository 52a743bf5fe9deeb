mkdir -p gpurun_out/s25
for IT in 1 2 3 4 6 0; do
  timeout 300 python bench.py --steps 300 --warmup 100 --no-cpu-baseline --no-e2e --iterations $IT > gpurun_out/s25/ch_it$IT.json 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s25/ch_it$IT.json") if l.startswith("{")][-1])
print("iterations $IT", "%.4f ms"%d["ms_per_step"], "%.1fM"%(d["value"]/1e6), "resident %.1fM"%(d.get("value_resident_one_launch",0)/1e6), {k:d.get(k) for k in ("mean_ncon","mean_nefc","mean_niter") if k in d})
PY
done
