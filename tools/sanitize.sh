#!/usr/bin/env bash
# compute-sanitizer over one case per kernel family.  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      -> gpurun_out/sanitizer_memcheck.txt
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh racecheck'     -> gpurun_out/sanitizer_racecheck.txt
# The same pytest selection runs once WITHOUT the tool first (it must pass) before the tool is attached.
set -uo pipefail
TOOL=${1:-memcheck}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='test_logmel_golden_and_port or (test_layernorm_bwd_single_pass and 1024-5003) or (test_layernorm_fwd_bwd and 301) or test_elementwise_and_reductions
 or (test_asp_pool_fwd_bwd and 500-1024) or (test_aam_softmax_vs_port and 7-13) or (test_arc_infonce_vs_port and dtype1) or (test_label_smoothed_ce and 1000)
 or (test_gemm_tcgen05_layouts and shape3) or test_gemm_epilogues or (test_gemm_column_sums_in_the_epilogue and shape1) or test_gemm_tcgen05_split_k_weight_gradient_shape
 or (test_gemm_skinny_weight_streaming and NK1-9) or test_fmha_fwd_moving_maximum or (test_fmha_fwd and 300-300) or (test_fmha_bwd and 300-300) or (test_fmha_bwd and 260-260) or (test_fmha_bwd and 1516-1516) or (test_fmha_fwd and 16-333)
 or (test_softmax_masks_fwd_bwd and dtype1) or test_decoder_embed_fwd_bwd'
SEL=$(echo $SEL)
FILES="tests/test_kernels_gpu.py"
python -m pytest $FILES -q -x -m gpu -k "$SEL" > gpurun_out/sanitizer_plain_${TOOL}.log 2>&1
rc=$?
tail -3 gpurun_out/sanitizer_plain_${TOOL}.log
if [ $rc -ne 0 ]; then echo "plain run failed (rc=$rc): not attaching compute-sanitizer"; exit $rc; fi
EXTRA=""
[ "$TOOL" = "memcheck" ] && EXTRA="--leak-check no --padding 32"
[ "$TOOL" = "racecheck" ] && EXTRA="--racecheck-report all"
timeout ${SANITIZE_TIMEOUT:-1200} compute-sanitizer --tool $TOOL $EXTRA --target-processes all --print-limit 50 --error-exitcode 0 \
  --log-file gpurun_out/sanitizer_${TOOL}.txt python -m pytest $FILES -q -x -m gpu -k "$SEL" > gpurun_out/sanitizer_${TOOL}_pytest.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
tail -3 gpurun_out/sanitizer_${TOOL}_pytest.log
tail -15 gpurun_out/sanitizer_${TOOL}.txt
