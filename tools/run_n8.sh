set -x
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/ulysses_multi_gpu_check.py > gpurun_out/sp_check_n$N.log 2>&1; echo "sp check rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/sp_check_n$N.log | tail -4
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err
