# A/B of the CTA-pair GEMM (tcgen05.mma cta_group::2, VAP_GEMM_PAIR=1) against the one-CTA kernel: parity checks, then TFLOP/s per shape.
set -x
VAP_GEMM_PAIR=1 timeout 150 python -m pytest tests -m gpu -x -q -k "gemm" 2>&1 | tail -5
for c in 0 1; do
  echo "== PAIR=$c"; VAP_GEMM_PAIR=$c timeout 200 python tools/kernel_bench.py --gemm 2>/dev/null | cut -c1-220
done
