# quick attention-kernel iteration on the GPU box: parity checks that touch attention, trace, micro-bench
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "attention or probe or native" > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/dev_pytest.log
timeout 120 python tools/attn_trace.py > gpurun_out/dev_trace128.log 2>&1; tail -14 gpurun_out/dev_trace128.log
timeout 120 python tools/attn_trace.py 64 > gpurun_out/dev_trace64.log 2>&1; tail -12 gpurun_out/dev_trace64.log
timeout 300 python tools/kernel_bench.py --attn --quick 2> gpurun_out/dev_kb.err | cut -c1-400
