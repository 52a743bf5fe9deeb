set -x
mkdir -p gpurun_out/s9
python tools/sanitize_run.py > gpurun_out/s9/plain.log 2>&1 &&
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/s9/r2_sanitizer_memcheck.log python tools/sanitize_run.py > gpurun_out/s9/memcheck_stdout.log 2>&1
echo "rc=$?" >> gpurun_out/s9/memcheck_stdout.log
tail -5 gpurun_out/s9/r2_sanitizer_memcheck.log; tail -3 gpurun_out/s9/memcheck_stdout.log
