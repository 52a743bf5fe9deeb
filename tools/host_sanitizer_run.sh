#!/bin/bash
# Runs the host instantiation of the product's stage templates (generic arena mode + every generated model specialisation,
# fp32 and fp64, zoo models included) under AddressSanitizer + UBSan and writes profiles/r2_host_asan_ubsan.log.
# compute-sanitizer is closed on the GPU pool (gpurun answers "closed on this pool"), so this is the bounds / UB evidence
# for the code the kernels share with the host instantiation; GPU-only code is covered by canary tests (tests/test_gpu_canary.py).
set -e
cd "$(dirname "$0")/.."
make -j8 -C tests/native CXX=/usr/bin/g++ libox_hostcheck_asan.so > /dev/null   # the system g++ ships libasan / libubsan
LOG=profiles/r2_host_asan_ubsan.log
ASAN=$(/usr/bin/gcc -print-file-name=libasan.so)
UBSAN=$(/usr/bin/gcc -print-file-name=libubsan.so)
{
  echo "# $(date -u +%FT%TZ) host instantiation under -fsanitize=address,undefined (g++ $(g++ -dumpversion)); LD_PRELOAD=$ASAN"
  OX_HOSTCHECK_SO=tests/native/libox_hostcheck_asan.so LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=0 \
    UBSAN_OPTIONS=print_stacktrace=1 python -m pytest tests/test_host_instantiation_parity.py tests/test_zoo_parity.py tests/test_n3_n4.py tests/test_touch.py \
    -q -m "not gpu" -p no:cacheprovider 2>&1 | grep -v "^$" | tail -25
} > "$LOG" 2>&1
echo "# libox_hostcheck_asan.so links: $(ldd tests/native/libox_hostcheck_asan.so | grep -o 'lib[a-z]*san\.so\.[0-9]*' | sort -u | tr '\n' ' ')" >> "$LOG"
echo "# sanitizer reports in this log: $(grep -c 'ERROR: AddressSanitizer\|runtime error:' "$LOG")" >> "$LOG"
tail -15 "$LOG"
