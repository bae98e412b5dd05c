#!/bin/bash
# GPU check of the training step: parity tests and the bench training pass (compute-sanitizer is closed on this pool).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --tb=short > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_train.log
timeout 300 python - <<'PY' > gpurun_out/train_pass.json 2> gpurun_out/train_pass.err
import json, bench
print(json.dumps(bench.train_step_pass(0, 5, 3)))
PY
echo "train pass rc=$?"; cat gpurun_out/train_pass.json; tail -3 gpurun_out/train_pass.err
