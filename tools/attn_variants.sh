for lib in build_variants/libvap_*.so; do
  VAP_B200_LIB=$PWD/$lib timeout 200 python tools/attn_variant_bench.py 2>&1 | tail -1
done
