# 2-GPU call: sharded context projections + batched peer exchange on real GPUs (both transports, Wan incl. B = 2, CogVideoX B = 2), then the bench.
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/ulysses_multi_gpu_check.py > gpurun_out/ulysses_check_wan_n2c.log 2>&1; echo "wan check rc=$?"; grep '^{' gpurun_out/ulysses_check_wan_n2c.log; tail -5 gpurun_out/ulysses_check_wan_n2c.log | cut -c1-300
timeout 300 $TR --master-port 29512 tools/ulysses_multi_gpu_check.py --family cog > gpurun_out/ulysses_check_cog_n2c.log 2>&1; echo "cog check rc=$?"; grep '^{' gpurun_out/ulysses_check_cog_n2c.log; tail -5 gpurun_out/ulysses_check_cog_n2c.log | cut -c1-300
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2c.log 2> gpurun_out/bench_n2c.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2c.log | cut -c1-2500; tail -3 gpurun_out/bench_n2c.err
