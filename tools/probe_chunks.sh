# D2H route probe over chunk sizes / ring depths (development; results in profiles/e2e_route_*.jsonl)
N=${1:-1}
for cfg in "4096 0" "1024 0" "512 16" "512 24" "256 16"; do set -- $cfg; echo "CHUNK_KB=$1 SLOTS=$2"
  if [ "$N" = 1 ]; then NIS_HOST_CHUNK_KB=$1 NIS_HOST_SLOTS=$2 python tools/e2e_route_probe.py 2>&1 | grep '"threads_per_rank": \(6\|8\|12\)' | cut -c1-250
  else NIS_HOST_CHUNK_KB=$1 NIS_HOST_SLOTS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/e2e_route_probe.py 2>&1 | grep '"threads_per_rank": \(3\|6\|8\)' | cut -c1-250; fi
done
