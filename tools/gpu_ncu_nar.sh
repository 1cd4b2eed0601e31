#!/bin/bash
# ncu --set full of one k_amp_nar launch pair (stage 5, k = 3, d = 1: layers A and B of the second decode)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout -s KILL 200 python tools/profile_decode.py 16 234 bf16 2 > gpurun_out/ncu_plain.log 2>&1 &&
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_amp_nar -s 90 -c 2 -o gpurun_out/r2c_nar_s5k3 -f python tools/profile_decode.py 16 234 bf16 2 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_plain.log; tail -5 gpurun_out/ncu_run.log; ls -la gpurun_out/*.ncu-rep | tail -2
