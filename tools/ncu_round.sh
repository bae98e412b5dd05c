# Round profile capture: one `ncu --set full` launch per major kernel class + launch lists.
# Usage on the GPU box: bash tools/ncu_round.sh <tag>   (writes gpurun_out/<tag>_*.ncu-rep / .csv)
# Under Nsight Compute kernels are serialised, so the c1 decode runs the single-kernel piped schedule
# (decode_band_pipe_kernel) instead of the wavefront pair.
TAG=${1:-r5}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
C1="python bench.py --steps 1 --warmup 1 --no-cpu"
C2="python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu"
$C1 > gpurun_out/plain_c1.log 2>&1 || exit 1
$NCU -k regex:decode_band_pipe -s 6 -c 1 -o gpurun_out/${TAG}_decode_pipe $C1 > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"
$NCU -k regex:cnn_tc_kernel -s 17 -c 1 -o gpurun_out/${TAG}_cnn_c1 $C1 > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
$NCU -k regex:encode_all_pair -s 2 -c 1 -o gpurun_out/${TAG}_encode_pair $C1 > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
$C2 > gpurun_out/plain_c2.log 2>&1 || exit 1
$NCU -k regex:cnn_tc_kernel -s 44 -c 1 -o gpurun_out/${TAG}_cnn_c2 $C2 > gpurun_out/ncu4.log 2>&1; echo "ncu4 rc=$?"
$NCU -k regex:window_kernel -s 39 -c 1 -o gpurun_out/${TAG}_window_c2 $C2 > gpurun_out/ncu5.log 2>&1; echo "ncu5 rc=$?"
$NCU -k regex:consume_kernel -s 39 -c 1 -o gpurun_out/${TAG}_consume_c2 $C2 > gpurun_out/ncu6.log 2>&1; echo "ncu6 rc=$?"
$NCU -k regex:band_bounds -s 27 -c 1 -o gpurun_out/${TAG}_bounds_c2 $C2 > gpurun_out/ncu6b.log 2>&1; echo "ncu6b rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_c1.csv $C1 > gpurun_out/ncu7.log 2>&1; echo "ncu7 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_c2.csv $C2 > gpurun_out/ncu8.log 2>&1; echo "ncu8 rc=$?"
ls -la gpurun_out/${TAG}_*
