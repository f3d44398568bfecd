# One development GPU pass: the whole parity suite on the default kernels, the 2-block Wan-14B-width step with and without the
# CTA-pair GEMM, then the pair-kernel tuning variants (tools/build_attn_variants.py with VAP_VARIANT_UNIT=gemm_sm100.cu).
set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in auto 0; do
  VAP_GEMM_PAIR=$c timeout 300 python bench.py --config wan14b_2l --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-330
done
echo "== in-tree"; timeout 100 python tools/kernel_bench.py --gemm --quick 2>/dev/null | cut -c100-330
for v in st6 g4 g8 g32; do
  echo "== $v"; VAP_B200_LIB=$PWD/build_variants/libvap_$v.so timeout 100 python tools/kernel_bench.py --gemm --quick 2>/dev/null | cut -c100-330
done
