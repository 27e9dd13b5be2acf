#!/bin/bash
# compute-sanitizer over the small GPU parity tests (every kernel, every frame size and mode, blocking and pipelined calls).
# Run on the GPU box:  gpurun --timeout 600 -- bash tools/sanitize_gpu.sh [memcheck|racecheck|initcheck|synccheck ...]
# Logs go to gpurun_out/sanitizer_<tool>.log; the tests still compare every sample with the reference while the tool watches.
mkdir -p gpurun_out
K_BIG="golden_fixture or golden_short or celt_layer_all_sizes or mode_switching or config5_mixed or tiny_and_maximal or filler_and_continuation or three_async or async_calls_in_flight or bad_frames or leaf_list_overflow or reset_subset"
K_SMALL="celt_layer_all_sizes or mode_switching or golden_short"
for tool in "${@:-memcheck}"; do
    k="$K_BIG"; t=200
    if [ "$tool" != memcheck ]; then k="$K_SMALL"; t=150; fi
    timeout $t compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 40 \
        python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$k" > gpurun_out/sanitizer_$tool.log 2>&1
    echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log) summaries: $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1) | $(tail -1 gpurun_out/sanitizer_$tool.log)"
done
