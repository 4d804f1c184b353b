#!/bin/bash
# Round-end measurement on ONE B200 (run under gpurun from the repo root):
#   bash tools/profile_round.sh r02
# 1. bench line (device-resident value, e2e, roofline, cpu_baseline, other_configs)  -> gpurun_out/<tag>_bench.json
# 2. reference arm (oracle port on the host cores)                                    -> gpurun_out/<tag>_bench_reference.json
# 3. ncu launch list of exactly one timed step (NVTX range wmk_timed_step)            -> gpurun_out/<tag>_launches.csv
# 4. one ncu --set full capture of the first 110 launches of a step (all stage widths C = 32..256 appear)
#                                                                                     -> gpurun_out/<tag>_top_raw.csv
# 5. stand-alone STFT / ISTFT / training-STFT throughput                              -> gpurun_out/<tag>_stft_bench.json
# Numbers printed by the runs under ncu are never bench values.  tools/make_profiles.py turns 3 and 4 into profiles/.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python tools/stft_bench.py > $OUT/${TAG}_stft_bench.json 2> $OUT/${TAG}_stft_bench.err
timeout 900 ncu --nvtx --nvtx-include "wmk_timed_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 1200 ncu --nvtx --nvtx-include "wmk_timed_step/" --set full --import-source on --clock-control none -c 110 \
    -o /tmp/${TAG}_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/${TAG}_ncu_top.log 2>&1
ncu -i /tmp/${TAG}_top.ncu-rep --page raw --csv > $OUT/${TAG}_top_raw.csv 2>/dev/null
# 6. launch list of the ModelA training step (BASELINE configs[4])                     -> gpurun_out/<tag>_launches_train.csv
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches_train.csv python tools/train_bench.py --steps 3 > $OUT/${TAG}_ncu_train.log 2>&1
ls -la $OUT | tail -8
