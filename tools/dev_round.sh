# One development GPU pass: the whole parity suite (default kernels, including the split-KV checks), then the CTA-pair GEMM
# (VAP_GEMM_PAIR=1): parity and TFLOP/s per shape against the one-CTA kernel and cuBLAS.
set -x
timeout 400 python -m pytest tests -m gpu -x -q -s -k "not gemm_large and not full_size" 2>&1 | grep -E "splitkv|passed|failed|Error|error" | cut -c1-400 | tail -20
timeout 200 python -m pytest tests -m gpu -x -q -s -k "attn_splitkv_sp8_shape or attn_full_size" 2>&1 | grep -E "splitkv|full_size|passed|failed|rror" | cut -c1-500 | tail
VAP_GEMM_PAIR=1 timeout 150 python -m pytest tests -m gpu -x -q -k "gemm" 2>&1 | tail -5
for c in 0 1; do
  echo "== PAIR=$c"; VAP_GEMM_PAIR=$c timeout 200 python tools/kernel_bench.py --gemm 2>/dev/null | cut -c1-220
done
