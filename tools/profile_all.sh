#!/bin/bash
# One GPU call that regenerates every artefact under gpurun_out/ that tools/make_profiles.py turns into profiles/:
# plain bench (must exit 0 first), ncu launch list of the same command, one --set full capture per dominant kernel.
set -u
TAG=${1:-r2}
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "bench failed"; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
NIS_BENCH_SKIP_MULTI=1 python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
NIS_BENCH_SKIP_MULTI=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
FULL="--set full --clock-control none --import-source on"
ncu $FULL -k regex:k_range -c 1 -o gpurun_out/prof_range_$TAG python tools/one_csa.py 8192 2 > /dev/null 2>&1
ncu $FULL -k regex:k_az_cluster -c 2 -o gpurun_out/prof_azc4096_$TAG python tools/one_csa.py 4096 2 > /dev/null 2>&1
ncu $FULL -k regex:k_az_inner -c 1 -o gpurun_out/prof_azinner_$TAG python tools/one_csa.py 8192 2 > /dev/null 2>&1
ncu $FULL -k regex:"k_az_outer" -c 2 -o gpurun_out/prof_azouter_$TAG python tools/one_csa.py 8192 1 > /dev/null 2>&1
ncu $FULL -k regex:k_echo -c 1 -o gpurun_out/prof_echo_sparse_$TAG python tools/kbench.py echo:stripmap8192 > /dev/null 2>&1
ncu $FULL -k regex:k_echo -c 1 -o gpurun_out/prof_echo_dense_$TAG python tools/kbench.py echo:ati_default_1184p > /dev/null 2>&1
ncu $FULL -k regex:"k_row_mixed_ct|k_row_blue_pruned" -c 3 -o gpurun_out/prof_general_$TAG python tools/one_csa.py 7199x13200 1 > /dev/null 2>&1
ncu $FULL -k regex:"k_gmti" -c 2 -o gpurun_out/prof_gmti_$TAG python tools/kbench.py gmti:4096 > /dev/null 2>&1
ls -la gpurun_out/*_$TAG.* | awk '{print $5, $9}'
# summarise on the box (the .ncu-rep files together exceed what a gpurun call brings back) and keep only the summaries
mkdir -p gpurun_out/profiles_$TAG
python tools/make_profiles.py $TAG --launches gpurun_out/launches_$TAG.csv --bench gpurun_out/bench_$TAG.json \
    --full gpurun_out/prof_range_$TAG.ncu-rep gpurun_out/prof_azc4096_$TAG.ncu-rep gpurun_out/prof_azinner_$TAG.ncu-rep \
           gpurun_out/prof_azouter_$TAG.ncu-rep gpurun_out/prof_echo_sparse_$TAG.ncu-rep gpurun_out/prof_echo_dense_$TAG.ncu-rep \
           gpurun_out/prof_general_$TAG.ncu-rep gpurun_out/prof_gmti_$TAG.ncu-rep > gpurun_out/make_profiles_$TAG.log 2>&1
cp profiles/*_$TAG.* profiles/traffic.json gpurun_out/profiles_$TAG/ 2>/dev/null
rm -f gpurun_out/*.ncu-rep
