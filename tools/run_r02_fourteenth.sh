set -x
timeout 300 python tools/norm_rows_ab.py > gpurun_out/norm_rows_ab.json 2> gpurun_out/norm_rows_ab.err; echo "norm_rows rc=$?"; cat gpurun_out/norm_rows_ab.json; tail -3 gpurun_out/norm_rows_ab.err
timeout 300 python tools/attn_sustained.py --seconds 8 > gpurun_out/attn_sustained.json 2> gpurun_out/attn_sustained.err; echo "sustained rc=$?"; cat gpurun_out/attn_sustained.json; tail -3 gpurun_out/attn_sustained.err
