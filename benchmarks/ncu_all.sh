#!/bin/bash
# ncu --set full capture of every config's dominant kernels (one GPU), summarised on the box so that only small
# files travel back (gpurun_out/ is capped at 64 MiB).  Each capture runs only after the plain command exited 0.
#   bash benchmarks/ncu_all.sh [configs...]      (default: C5 C4 C3 C2 M MM)
mkdir -p gpurun_out
CFGS="${@:-C5 C4 C3 C2 M MM}"
R=${ROUND:-r02}
for cfg in $CFGS; do
  python benchmarks/prof_one.py $cfg > gpurun_out/plain_$cfg.log 2>&1 || { echo "plain $cfg failed"; tail -3 gpurun_out/plain_$cfg.log; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'chain3|dmma_kernel|dmma_ws_kernel|dmma_wa_kernel|tridiag_solve|kron_ring|tridiag_ring|tridiag_dot|gbmv_tile|gbmv_multi|gbmv2|gbmm_reg|tridiag_axpby|tridiag_mm|tridiag_mv' \
    -o gpurun_out/prof_${cfg}_${R} -f python benchmarks/prof_one.py $cfg > gpurun_out/ncu_${cfg}_e.log 2>&1
  tail -1 gpurun_out/ncu_${cfg}_e.log
  python benchmarks/ncu_summary.py gpurun_out/prof_${cfg}_${R}.ncu-rep gpurun_out/ncu_${cfg}_${R}.md --last-per-kernel
  case $cfg in KEEP) ;; *) rm -f gpurun_out/prof_${cfg}_${R}.ncu-rep ;; esac
done
du -sh gpurun_out
