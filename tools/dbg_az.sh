for v in 0 1 2 3 4; do echo "OUTER_INV=$v"; NIS_OUTER_INV=$v python tools/sched_bench.py azone 8192 0 2>&1 | tail -1 | cut -c150-290; done
