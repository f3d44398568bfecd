set -x
VAP_GEMM_MT=2 timeout 300 python -m pytest tests -m gpu -x -q -k "gemm or wan_blocks or cog_blocks" 2>&1 | tail -2
for mt in 1 2; do
  echo "== MT=$mt"; VAP_GEMM_MT=$mt timeout 300 python tools/kernel_bench.py --gemm 2>/dev/null | cut -c1-200
done
for mt in 1 2; do
  VAP_GEMM_MT=$mt timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-220
done
