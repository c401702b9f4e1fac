#!/bin/bash
# compute-sanitizer pass (memcheck, racecheck, synccheck, initcheck) over a small set of GPU tests that covers every kernel family:
# the trust-region solve + marginalisation, the tracker sequence (CLAHE, pyramid, LK, corners, set_mask, RANSAC) and the
# initialisation, the device FeatureManager and the solve from its list.  Run on a GPU box:   tools/sanitize.sh [out_dir]      (summaries go to out_dir, default gpurun_out/)
# Each tool runs under its own timeout; a tool that does not finish is reported as such, not as clean.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
TESTS="tests/test_backend_gpu.py::test_optimize_matches_oracle tests/test_backend_gpu.py::test_marginalize_chain_with_prior_and_second_new tests/test_frontend_gpu.py::test_tracker_sequence_matches_golden tests/test_ransac_gpu.py tests/test_initial_gpu.py::test_global_sfm tests/test_prestep_gpu.py tests/test_feature_store_gpu.py::test_feature_store_stream tests/test_feature_store_gpu.py::test_device_store_three_steps"
SUMMARY="$OUT/r2_sanitizer_summary.txt"
: > "$SUMMARY"
for tool in memcheck racecheck synccheck initcheck; do
    log="$OUT/r2_sanitizer_$tool.log"
    extra=""
    [ "$tool" = racecheck ] && extra="--racecheck-report analysis"
    [ "$tool" = memcheck ] && extra="--leak-check no"
    start=$(date +%s)
    timeout ${SAN_TIMEOUT:-900} "$SAN" --tool $tool $extra --target-processes all --print-limit 50 \
        python -m pytest $TESTS -m gpu -x -q -p no:cacheprovider > "$log" 2>&1
    rc=$?
    end=$(date +%s)
    errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -3 | tr '\n' ' ')
    tests_line=$(grep -E "passed|failed|error" "$log" | tail -1)
    echo "$tool: rc=$rc wall=$((end - start))s | $tests_line | $errs" >> "$SUMMARY"
    grep -E "========= (Error|Warning|Race|Barrier|Uninitialized|Invalid|Hazard)" "$log" | sort | uniq -c | sort -rn | head -20 >> "$SUMMARY"
done
cat "$SUMMARY"
