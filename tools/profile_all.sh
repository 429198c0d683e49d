#!/bin/bash
# One GPU call that regenerates every artefact under gpurun_out/ that tools/make_profiles.py turns into profiles/:
# plain bench (must exit 0 first), ncu launch list of the same command, one --set full capture per dominant kernel.
set -u
TAG=${1:-r1}
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "bench failed"; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
FULL="--set full --clock-control none --import-source on"
ncu $FULL -k regex:k_range -c 1 -o gpurun_out/prof_range_$TAG python tools/one_csa.py 8192 2 > /dev/null 2>&1
ncu $FULL -k regex:k_az_cluster -c 2 -o gpurun_out/prof_azc4096_$TAG python tools/one_csa.py 4096 2 > /dev/null 2>&1
ncu $FULL -k regex:k_az_inner -c 1 -o gpurun_out/prof_azinner_$TAG python tools/one_csa.py 8192 2 > /dev/null 2>&1
ncu $FULL -k regex:k_echo -c 1 -o gpurun_out/prof_echo_$TAG python tools/kbench.py echo:stripmap8192 > /dev/null 2>&1
ncu $FULL -k regex:"k_rda_range|k_rda_rcmc" -c 2 -o gpurun_out/prof_rda_$TAG python tools/kbench.py rda:4096x4096 > /dev/null 2>&1
ncu $FULL -k regex:"k_tdbp" -c 3 -o gpurun_out/prof_tdbp_$TAG python tools/kbench.py tdbp:2500x512 > /dev/null 2>&1
ncu $FULL -k regex:"k_row_mixed_ct|k_row_blue_pruned" -c 3 -o gpurun_out/prof_general_$TAG python tools/one_csa.py 7199x13200 1 > /dev/null 2>&1
ncu $FULL -k regex:"k_gmti_products" -c 1 -o gpurun_out/prof_gmti_$TAG python tools/kbench.py gmti:4096 > /dev/null 2>&1
ls -la gpurun_out/*_$TAG.* | awk '{print $5, $9}'
