#!/bin/bash
# Roofline evidence at the current commit: per-launch table (CUDA events), ncu launch list (time + DRAM bytes) and
# ncu --set full captures of representative launches.  Usage: bash tools/gpu_evidence.sh <tag>
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${1:-r02}
P="python tools/profile_decode.py 16 234 bf16 2"
BVG_PROF_DUMP=1 timeout -s KILL 200 python tools/per_launch.py 2> gpurun_out/${TAG}_per_launch.txt
timeout -s KILL 200 $P > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:k_amp_tc -s 115 -c 115 --csv --log-file gpurun_out/${TAG}_launches_time_dram.csv $P > /dev/null 2>&1
cap() {  # name, kernel regex, skip, tiles per launch (for the per-role table).  The .ncu-rep files are summarised here and
  # deleted (gpurun brings back at most 64 MiB); KEEP_REP=1 keeps them
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 $P > gpurun_out/${TAG}_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_$1.ncu-rep > gpurun_out/${TAG}_ncu_$1.txt 2>&1
  python tools/ncu_roles.py gpurun_out/${TAG}_$1.ncu-rep $4 >> gpurun_out/${TAG}_ncu_$1.txt 2>&1
  [ "$KEEP_REP" = "1" ] || rm -f gpurun_out/${TAG}_$1.ncu-rep
}
# tiles per launch: 16 utterances x ceil(234 * rate / 256) time tiles x column tiles
cap s0k3A k_amp_tc 117 192
cap s0k7A k_amp_tc 123 192
cap s0k11A k_amp_tc 129 192
cap s1k3A k_amp_tc 136 480
cap s3k7A k_amp_tc 180 3744
cap s5k3B k_amp_tc 213 14976
cap s5k11B k_amp_tc 229 14976
cap actblk k_act_blk 1 1
ls gpurun_out/${TAG}_ncu_*.txt | wc -l; du -sh gpurun_out
