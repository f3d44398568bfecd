set -x
python tools/prof_target.py attn > gpurun_out/plain_attn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/attn_dev -f python tools/prof_target.py attn > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
