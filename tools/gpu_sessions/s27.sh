mkdir -p gpurun_out/s27
for rep in 1 2; do for st in 1 0; do for sc in 1 0.125; do
OX_B200_COOP_STATIC=$st timeout 300 python bench.py --config humanoid --steps 300 --warmup 100 --no-cpu-baseline --no-e2e --ctrl-scale $sc > gpurun_out/s27/hum_st${st}_s$sc.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s27/hum_st${st}_s$sc.json") if l.startswith("{")][-1])
print("rep $rep static $st scale $sc", "%.4f ms"%d["ms_per_step"], "%.2fM"%(d["value"]/1e6), "resident %.2fM"%(d.get("value_resident_one_launch",0)/1e6))
PY
done; done; done
CMD="python tools/prof_step.py --config humanoid --nenv 4096 --launches 10 --warmup 100"
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active --clock-control none -k regex:k_solve_coop -s 100 -c 5 --csv --log-file gpurun_out/s27/solve_dynamic.csv $CMD > /dev/null 2>&1
OX_B200_COOP_STATIC=1 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active --clock-control none -k regex:k_solve_coop -s 100 -c 5 --csv --log-file gpurun_out/s27/solve_static.csv $CMD > /dev/null 2>&1
tail -n 15 gpurun_out/s27/solve_dynamic.csv | cut -d, -f 5,12- ; tail -n 15 gpurun_out/s27/solve_static.csv | cut -d, -f 5,12-
