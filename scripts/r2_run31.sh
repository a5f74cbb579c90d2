#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: token path (attention rows, batched GEMM, softmax, LayerNorm), the
# (test_attention_tc_matches_library is left out: synccheck reports a "missing init" barrier inside cuDNN's fused-attention kernel,
# the library path that test compares against — not a kernel of this repository)
# double-buffered small-channel weight gradient, channel_stats, the 128-channel producer budget
mkdir -p gpurun_out
for TOOL in ${TOOLS:-memcheck synccheck}; do
OUT=gpurun_out/sanitize2_$TOOL.log
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 77 --launch-timeout 120 \
    python -m pytest tests/test_token_path_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider \
    -k "(attention or layer_norm or c2+0_2 or c4+0_4 or channel_stats or c64+64_64_k3 or c128+0_64_k3) and not matches_library" > $OUT 2>&1
echo "compute-sanitizer --tool $TOOL rc=$? (77 = errors reported)"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|Uninitialized|hazard" $OUT | tail -6
done
