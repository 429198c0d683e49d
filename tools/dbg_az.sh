for w in 8 16 4; do echo "W=$w"; NIS_INNER_W=$w python tools/sched_bench.py azone 8192 0 2>&1 | tail -1 | cut -c1-290; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
