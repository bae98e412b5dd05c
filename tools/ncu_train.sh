#!/bin/bash
# ncu captures of the training step's kernels (scale 0, band 2 of the reference's training batch) + launch list.
mkdir -p gpurun_out
cat > /tmp/train_pass.py <<PY
import sys; sys.path.insert(0, "$PWD"); import json, bench
print(json.dumps(bench.train_step_pass(0, 1, 1)))
PY
timeout 300 python /tmp/train_pass.py > gpurun_out/train_plain.json 2> gpurun_out/train_plain.err || exit 1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:cnn_backward -s 2 -c 1 -o gpurun_out/r02_cnn_backward python /tmp/train_pass.py > gpurun_out/ncu_t1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 $NCU -k regex:cnn_forward_train -s 2 -c 1 -o gpurun_out/r02_cnn_forward_train python /tmp/train_pass.py > gpurun_out/ncu_t2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 $NCU -k regex:self_info_grad -s 2 -c 1 -o gpurun_out/r02_self_info_grad python /tmp/train_pass.py > gpurun_out/ncu_t3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/r02_cnn_backward*
