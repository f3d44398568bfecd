# 2 real GPUs, split-KV attention FORCED on (the path 8-way Ulysses takes at 480p): parity of both transports against a single-GPU
# forward, then the 2-block step.
set -x
VAP_ATTN_SPLITKV=2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/ulysses_multi_gpu_check.py > gpurun_out/sp_check_splitkv.log 2>&1; echo "sp check rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/sp_check_splitkv.log | tail -6
for s in 2 auto; do
VAP_ATTN_SPLITKV=$s timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config wan14b_2l --steps 4 --warmup 3 > gpurun_out/bench_n2_splitkv_$s.log 2> gpurun_out/bench_n2_splitkv_$s.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n2_splitkv_$s.log | cut -c1-300; tail -3 gpurun_out/bench_n2_splitkv_$s.err
done
