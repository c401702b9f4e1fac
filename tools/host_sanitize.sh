#!/bin/bash
# Builds liblfvio_host.so with AddressSanitizer + UBSan into a scratch copy and runs the CPU-only host tests against it
# (FeatureManager bookkeeping, node message synchronisation, gauge re-anchoring, C-ABI export check).  The product library is
# restored afterwards.  Usage, from the repo root:  bash tools/host_sanitize.sh
set -e
cd "$(dirname "$0")/.."
cp lf-vio_b200/liblfvio_host.so /tmp/liblfvio_host.so.bak
trap 'cp /tmp/liblfvio_host.so.bak lf-vio_b200/liblfvio_host.so' EXIT
(cd lf-vio_b200 && g++ -O1 -g -std=c++17 -fPIC -Wall -fsanitize=address,undefined -fno-omit-frame-pointer -I../include -shared \
    -o liblfvio_host.so host/*.cpp -L. -llfvio -Wl,-rpath,'$ORIGIN')
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    python -m pytest tests/test_feature_manager_host.py tests/test_nodes_host.py tests/test_gauge_fix_host.py tests/test_abi.py -q -x
