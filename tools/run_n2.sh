set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "ulysses or attn or qk" > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/ulysses_multi_gpu_check.py > gpurun_out/sp_check.log 2>&1; echo "sp check rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/sp_check.log | tail -8
for mode in p2p nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --sp-mode $mode > gpurun_out/bench_n2_$mode.log 2> gpurun_out/bench_n2_$mode.err; echo "bench $mode rc=$?"; tail -1 gpurun_out/bench_n2_$mode.log | cut -c1-260; tail -3 gpurun_out/bench_n2_$mode.err
done
