mkdir -p gpurun_out/s31
for NE in 4736 8192 9472; do for BL in 32 64 128; do
  timeout 300 python bench.py --steps 300 --warmup 100 --no-cpu-baseline --no-e2e --nenv $NE --block $BL > gpurun_out/s31/ch_n${NE}_b$BL.json 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s31/ch_n${NE}_b$BL.json") if l.startswith("{")][-1])
print("nenv $NE block $BL", "%.4f ms"%d["ms_per_step"], "%.1fM"%(d["value"]/1e6), "resident %.1fM"%(d.get("value_resident_one_launch",0)/1e6))
PY
done; done
