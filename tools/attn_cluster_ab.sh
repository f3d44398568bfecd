set -x
VAP_ATTN_CLUSTER=2 timeout 200 python -m pytest tests -m gpu -x -q -k "attn or attention or ulysses_p2p or wan_blocks" 2>&1 | tail -3
for c in 0 2 0 2; do
  VAP_ATTN_CLUSTER=$c timeout 200 python tools/attn_variant_bench.py 2>&1 | tail -1
done
for c in 0 2; do
  VAP_ATTN_CLUSTER=$c timeout 400 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks'])"
done
