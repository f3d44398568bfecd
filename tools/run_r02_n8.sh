# Round 2, 8-GPU call: the headline workload sharded 8 ways — eager pass and CUDA-graph replay in ONE run (bench.py --graph on reports both), with parity_vs_single.
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --graph on > gpurun_out/bench_n8_graph.log 2> gpurun_out/bench_n8_graph.err; echo "bench n8 graph rc=$?"; grep '^{' gpurun_out/bench_n8_graph.log | cut -c1-1800; tail -3 gpurun_out/bench_n8_graph.err | cut -c1-300
