# 2 real GPUs: CogVideoX-VAP under Ulysses (both transports, B = 2) against a single-GPU forward; Wan again as the control.
set -x
for fam in cog; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/ulysses_multi_gpu_check.py --family $fam > gpurun_out/sp_check_$fam.log 2>&1; echo "sp check $fam rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/sp_check_$fam.log | tail -8
done
