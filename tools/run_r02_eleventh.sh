# launch list of the per-rank proxy (2 MoT blocks at Wan-14B widths, 2 535 rows per stream): where a rank's non-attention time goes under 8-way Ulysses
set -x
C="python tools/dual_stream_ab.py --case rank_of_8 --profile"
$C > gpurun_out/proxy_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rank_proxy.csv $C > gpurun_out/ncu_proxy.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_proxy.log; wc -l gpurun_out/launches_rank_proxy.csv
