#!/bin/bash
# Round-end measurement on ONE B200 (run under gpurun from the repo root):
#   bash tools/profile_round.sh r01
# 1. bench line (device-resident value, e2e, roofline, cpu_baseline)           -> gpurun_out/<tag>_bench.json
# 2. ncu launch list of exactly one timed step (NVTX range wmk_timed_step)      -> gpurun_out/<tag>_launches.csv
# 3. one ncu --set full capture of the first 130 launches of a step (all stage widths C = 32..256 appear)
#                                                                              -> gpurun_out/<tag>_top_raw.csv
# 4. stand-alone STFT / ISTFT / training-STFT throughput                        -> gpurun_out/<tag>_stft_bench.json
# Numbers printed by the runs under ncu are never bench values.  tools/make_profiles.py turns 2 and 3 into profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || exit 1
python tools/stft_bench.py > $OUT/${TAG}_stft_bench.json 2> $OUT/${TAG}_stft_bench.err
timeout 900 ncu --nvtx --nvtx-include "wmk_timed_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 1200 ncu --nvtx --nvtx-include "wmk_timed_step/" --set full --import-source on --clock-control none -c 130 \
    -o /tmp/${TAG}_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_top.log 2>&1
ncu -i /tmp/${TAG}_top.ncu-rep --page raw --csv > $OUT/${TAG}_top_raw.csv 2>/dev/null
ls -la $OUT | tail -8
