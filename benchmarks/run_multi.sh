#!/bin/bash
# N-GPU parity + benchmarks (one box): bash benchmarks/run_multi.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 600 $TR benchmarks/check_shard_nccl.py > gpurun_out/shard_n$N.log 2>&1; echo "check rc=$?"; tail -2 gpurun_out/shard_n$N.log
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_peer.json 2> gpurun_out/bench_n${N}_peer.err; echo "bench peer rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --nccl-exchange > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "bench nccl rc=$?"
timeout 900 $TR benchmarks/configs_multi.py > gpurun_out/configs_multi_n$N.jsonl 2> gpurun_out/configs_multi_n$N.err; echo "configs rc=$?"
for f in gpurun_out/bench_n${N}_peer.json gpurun_out/bench_n${N}_nccl.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], d["frac_of_aggregate_peak"], d["config"]["exchange"][:40])
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
cut -c1-200 gpurun_out/configs_multi_n$N.jsonl
