set -x
VAP_GEMM_CLUSTER=2 timeout 120 python -m pytest tests -m gpu -x -q -k "gemm" 2>&1 | tail -3
VAP_GEMM_CLUSTER=2 timeout 200 python -m pytest tests -m gpu -x -q -k "wan_blocks or cog_blocks or wan_model" 2>&1 | tail -2
for c in 0 2; do
  echo "== CLUSTER=$c"; VAP_GEMM_CLUSTER=$c timeout 300 python tools/kernel_bench.py --gemm 2>/dev/null | cut -c1-200
done
for c in 0 2; do
  VAP_GEMM_CLUSTER=$c timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-220
done
