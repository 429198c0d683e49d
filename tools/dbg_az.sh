for sk in 0 500 1000 2000 3000 4000 6000; do echo "skew $sk"; NIS_RANGE_SKEW=$sk python tools/sched_bench.py azone 4096 0 2>&1 | tail -1 | cut -c1-250; done
for sk in 0 500 1000 2000; do echo "skew $sk"; NIS_RANGE_SKEW=$sk python tools/sched_bench.py azone 2048 0 2>&1 | tail -1 | cut -c1-250; done
