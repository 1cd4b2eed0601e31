#!/bin/bash
# tests + bench at the current commit.  Usage: bash tools/gpu_call3.sh <tag>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${1:-r2x}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
timeout -s KILL 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 1500 gpurun_out/${TAG}_bench.json
