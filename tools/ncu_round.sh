mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain_c1.log 2>&1 || exit 1
$NCU -k regex:decode_band_pipe -s 9 -c 1 -o gpurun_out/r3_decode_pipe python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"
$NCU -k regex:cnn_tc_kernel -s 17 -c 1 -o gpurun_out/r3_cnn_c1 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
$NCU -k regex:encode_all_warp -s 2 -c 1 -o gpurun_out/r3_encode_warp python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/plain_c2.log 2>&1 || exit 1
$NCU -k regex:cnn_tc_kernel -s 44 -c 1 -o gpurun_out/r3_cnn_c2 python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu4.log 2>&1; echo "ncu4 rc=$?"
$NCU -k regex:window_kernel -s 39 -c 1 -o gpurun_out/r3_window_c2 python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu5.log 2>&1; echo "ncu5 rc=$?"
$NCU -k regex:consume_kernel -s 39 -c 1 -o gpurun_out/r3_consume_c2 python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu6.log 2>&1; echo "ncu6 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu7.log 2>&1; echo "ncu7 rc=$?"
