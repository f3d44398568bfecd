# Round 2, fourth GPU call: clock64 timeline of one CTA of the joint attention (16-lane kernel) + A/B of busy-polling waits on the critical chain.
set -x
VAP_B200_LIB=$PWD/build_variants/libvap_l16trace.so timeout 120 python tools/attn_trace.py 128 > gpurun_out/attn_trace_d128.txt 2>&1; tail -40 gpurun_out/attn_trace_d128.txt
VAP_B200_LIB=$PWD/build_variants/libvap_l16trace.so timeout 120 python tools/attn_trace.py 64 > gpurun_out/attn_trace_d64.txt 2>&1; tail -12 gpurun_out/attn_trace_d64.txt
timeout 400 python tools/attn_ab.py --rounds 3 --shapes wan,cog --only intree,l16spin1,l16spin2,l16spin3,rowp1f0,rowp1f0spin3 > gpurun_out/attn_ab2.json 2> gpurun_out/attn_ab2.err; echo "attn_ab rc=$?"; tail -c 2500 gpurun_out/attn_ab2.json; tail -5 gpurun_out/attn_ab2.err
