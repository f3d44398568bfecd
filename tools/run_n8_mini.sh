# 8 GPUs, short: the 2-block Wan-14B-width step under 8-way Ulysses (5 heads per rank: the split-KV attention switches itself on).
set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --config wan14b_2l --steps 4 --warmup 3 > gpurun_out/bench_n8_2l.log 2> gpurun_out/bench_n8_2l.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n8_2l.log | cut -c1-400; tail -3 gpurun_out/bench_n8_2l.err
