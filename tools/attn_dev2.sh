set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "attention or probe or native" > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/dev_pytest.log
bash tools/attn_variants.sh
VAP_B200_LIB=$PWD/build_variants/libvap_trace.so timeout 120 python tools/attn_trace.py > gpurun_out/dev_trace128.log 2>&1; tail -12 gpurun_out/dev_trace128.log
