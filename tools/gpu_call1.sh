#!/bin/bash
# first GPU call of round 2: tests, smoke, bench, per-launch table, k_amp_fir A/B
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
BVG_PROF_DUMP=1 python tools/per_launch.py 2> gpurun_out/r2a_per_launch.txt
python tools/time_layers.py > gpurun_out/r2a_time_layers_default.txt 2>&1
BVG_FIR_MAX_C=48 python tools/time_layers.py > gpurun_out/r2a_time_layers_fir48.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
tail -5 gpurun_out/r2a_pytest.txt; tail -3 gpurun_out/r2a_smoke.txt; cat gpurun_out/r2a_time_layers_*.txt | grep BVG_DBG; head -c 600 gpurun_out/r2a_bench.json
