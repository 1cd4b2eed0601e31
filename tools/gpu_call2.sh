#!/bin/bash
# k_amp_nar bring-up: op-level parity first (a hang must not take the box down), then generator, then timing A/B
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -k "amp_layer_bf16_tcgen05" 2>&1 | tail -25 > gpurun_out/r2b_op.txt
tail -8 gpurun_out/r2b_op.txt
if grep -q "passed" gpurun_out/r2b_op.txt && ! grep -q "failed" gpurun_out/r2b_op.txt; then
  timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.txt
  tail -5 gpurun_out/r2b_pytest.txt
  timeout -s KILL 200 python tools/time_layers.py > gpurun_out/r2b_time_layers_nar.txt 2>&1
  BVG_NAR_MAX_C=0 timeout -s KILL 200 python tools/time_layers.py > gpurun_out/r2b_time_layers_tc.txt 2>&1
  BVG_PROF_DUMP=1 timeout -s KILL 200 python tools/per_launch.py 2> gpurun_out/r2b_per_launch.txt
  cat gpurun_out/r2b_time_layers_*.txt | grep BVG_DBG
fi
