#!/bin/bash
# Round-end multi-GPU evidence on one box: bash benchmarks/final_multi.sh N OUTDIR
N=${1:-2}
OUT=${2:-gpurun_out/final_n$N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"
timeout 400 $TR benchmarks/configs_multi.py --only M,C2,C3 > $OUT/configs_multi_n$N.jsonl 2> $OUT/configs_multi_n$N.err; echo "configs rc=$?"
timeout 400 $TR benchmarks/check_shard_nccl.py > $OUT/shard_n$N.log 2>&1; echo "shard check rc=$?"; tail -2 $OUT/shard_n$N.log
timeout 300 python benchmarks/check_multi_single_process.py > $OUT/multi_single_process_n$N.log 2>&1; echo "single-process rc=$?"; tail -2 $OUT/multi_single_process_n$N.log
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3; fi
python - $OUT/bench_n$N.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d.get("frac_of_aggregate_peak"), d["parity"]["ok"], d["e2e"]["value"])
        for c in d["roofline"]["configs"]: print(c.get("config","")[:50], c.get("ms"), c.get("frac"), c.get("error"))
PY
cut -c1-230 $OUT/configs_multi_n$N.jsonl
