set -x
mkdir -p gpurun_out/s4
timeout 900 python tools/fp32_study.py > gpurun_out/s4/fp32_study.jsonl 2> gpurun_out/s4/fp32_study.err
cat gpurun_out/s4/fp32_study.jsonl; tail -5 gpurun_out/s4/fp32_study.err
