#!/bin/bash
# Does sustained FP64 DMMA load pull the SM clock down?  Samples nvidia-smi while (a) the register-only DMMA loop and
# (b) the banded product run back to back for ~2 s each.
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/fp64_clocks.csv &
SMI=$!
sleep 0.5
echo "phase idle $(date +%s.%N)" >> gpurun_out/fp64_phases.txt
benchmarks/micro/fp64_pipes long > gpurun_out/fp64_long.json
echo "phase microbench_done $(date +%s.%N)" >> gpurun_out/fp64_phases.txt
python - <<'PY'
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/lazybandedmatrices.jl_b200")
import torch, lbm_b200 as L
n = 1 << 20
A = L.brand(n, n, 128, 128, seed=1); B = L.brand(n, n, 128, 128, seed=2)
Cm = L.BandedMatrix(torch.empty((n, 513), dtype=torch.float64, device="cuda"), n, 256, 256)
h = L.handle_for(0)
for v in (22, 0):
    h.set_tuning(gbmm_variant=v)
    for reps in (3, 300):
        L.gbmm(A, B, out=Cm); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): L.gbmm(A, B, out=Cm)
        e1.record(); torch.cuda.synchronize()
        print(f"variant {v} reps {reps}: {e0.elapsed_time(e1) / reps:.3f} ms per product", flush=True)
PY
kill $SMI
sort -u gpurun_out/fp64_clocks.csv | sort -n | awk -F, '{print $1, $3, $4}' | sort | uniq -c | sort -rn | head -12
cat gpurun_out/fp64_long.json
