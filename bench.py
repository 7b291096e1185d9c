#!/usr/bin/env python
"""bench.py -- banded mul! HBM GB/s & % roofline (BASELINE.json metric).

One "step" = one pass of the hot path over the metric's single-GPU-sized workload:
    mul!(y, A, x) for A = brand(n,n,l,u) Float64, n = 2^27, (l,u) in {(1,1), (8,8)}
row-sharded over the N ranks (one process per GPU; ghost-cell halo exchange of x over NCCL
inside the timed region).  `value` = algorithmic bytes of both applies / step time, aggregated
over all ranks (GB/s).  The (128,128) point of the metric does not fit one GPU at n = 2^27
(A = 276 GB); it is reported in `breakdown` at the largest n that fits.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun ... bench.py --gpus N ...          (N > 1)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

BANDS = [(1, 1), (8, 8)]
LOG2N = 27
METRIC = "banded mul! aggregate HBM throughput (algorithmic bytes / time), n=2^27, l=u in {1,8}"


def alg_bytes(n, l, u, w=8):
    """gbmv (beta=0) moves w*[(l+u+1)*n + n + n] algorithmic bytes."""
    return w * n * (l + u + 1 + 2)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every ~2 ms from a
    thread (nvidia-smi's loop is too slow for a 40 ms region); falls back to nvidia-smi."""

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.mx = [], set(), None
        self._stop = threading.Event()
        self.t = None
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            hdl = None
            if uuid:
                for cand in ("GPU-" + uuid, uuid):
                    try:
                        hdl = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        hdl = None
            if hdl is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
                hdl = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(hdl, pynvml.NVML_CLOCK_SM))
            R = pynvml
            names = {"hw_slowdown": R.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": R.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": R.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": R.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM)))
                        m = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hdl)
                        for k, bit in names.items():
                            if m & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
            self.how = "nvml"
        except Exception:
            self.how = "nvidia-smi"
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                              "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)

                def rd():
                    for line in self.proc.stdout:
                        f = [t.strip() for t in line.split(",")]
                        try:
                            self.sm.append(float(f[0]))
                            self.mx = float(f[1])
                        except Exception:
                            continue
                        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                            if v.lower().startswith("active"):
                                self.reasons.add(name)
                self.t = threading.Thread(target=rd, daemon=True)
                self.t.start()
            except Exception:
                self.how = None

    def stop(self):
        self._stop.set()
        if self.how == "nvidia-smi":
            time.sleep(0.05)
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=1.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": self.how}


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: OpenBLAS ?gbmv -- the BLAS routine the reference's gbmv! ccalls
# --------------------------------------------------------------------------------------------
class CpuGbmvSample:
    """The reference's CPU path for this workload on the box's host cores: OpenBLAS ?gbmv through scipy -- the BLAS
    routine BandedMatrices.gbmv! ccalls -- with every host thread, on a bounded sample (same bands, reduced n);
    the repo's OpenMP oracle port is timed beside it.  Inputs are generated once; `step()` times one pass."""

    def __init__(self, log2n_sample):
        import numpy as np
        from oracle import oracle as O
        self.n = 1 << log2n_sample
        self.log2n = log2n_sample
        self.ops = []
        for (l, u) in BANDS:
            self.ops.append((l, u, O.brand(self.n, self.n, l, u, seed=1), O.fill_uniform(9, self.n), np.zeros(self.n)))
        self.bytes = sum(alg_bytes(self.n, l, u) for (l, u) in BANDS)
        self.cores = os.cpu_count() or 1
        try:
            from threadpoolctl import threadpool_info
            self.blas_threads = max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
        except Exception:
            self.blas_threads = None
        self.omp_threads = O.num_threads()

    def step(self):
        from scipy.linalg import blas
        from oracle import oracle as O
        tb = to = 0.0
        for (l, u, A, x, y) in self.ops:
            t0 = time.perf_counter()
            blas.dgbmv(self.n, self.n, l, u, 1.0, A, x, beta=0.0, y=y, overwrite_y=1)
            tb += time.perf_counter() - t0
            t0 = time.perf_counter()
            O.gbmv("N", self.n, self.n, l, u, 1.0, A, x)
            to += time.perf_counter() - t0
        return tb, to

    def describe(self, reps):
        return (f"dgbmv f64 n=2^{self.log2n}, l=u in {{1,8}} (same workload at reduced n; {reps} passes; "
                f"OpenBLAS threads={self.blas_threads}, OpenMP threads={self.omp_threads})")


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use all host threads."""
    n = str(os.cpu_count() or 1)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    _use_all_host_threads()          # before numpy / scipy / libgomp are loaded
    sample = CpuGbmvSample(24)
    for _ in range(max(1, args.warmup)):
        sample.step()
    tb = to = 0.0
    K = max(1, args.steps)
    for _ in range(K):
        a, b = sample.step()
        tb += a
        to += b
    v = sample.bytes * K / tb / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tb / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "gbmv f64 l=u in {1,8}, bounded sample n=2^24 of the n=2^27 workload",
                   "note": "reference is pure Julia (not runnable here); timed = OpenBLAS dgbmv (scipy.linalg.blas, "
                           "all host threads): the BLAS routine BandedMatrices.gbmv! ccalls, i.e. what the reference "
                           "executes for this path (BASELINE.md section 4). The repo's own OpenMP oracle port is "
                           "timed beside it as cpu_baseline.oracle_omp_gbs."},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": sample.cores, "kind": "port", "sample": sample.describe(K),
                         "which": "openblas_dgbmv", "oracle_omp_gbs": sample.bytes * K / to / 1e9},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import lbm_b200 as L
    from lbm_b200.shard import ShardedBanded, init_comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = L.handle_for(local)
    peer_on = False
    if world > 1 and not args.torch_p2p:
        # ghost exchange inside the library: one C call per sharded mul!; peer-memory mailboxes unless --nccl-exchange
        init_comm(local, peer=not args.nccl_exchange)
        peer_on = bool(L._lib.lib().lbm_peer_active(h.ptr))
    n = 1 << (args.log2n or LOG2N)
    K, W = args.steps, max(3, args.warmup)

    # ---- resident operands (generated on device from the counter-based stream) -------------
    shards, xbufs, ys = [], [], []
    for (l, u) in BANDS:
        sh = ShardedBanded.brand(n, l, u, rank, world, dtype=torch.float64, device=dev, seed=0x1)
        xb = sh.plan.new_vector(torch.float64, dev)
        own = sh.plan.owned(xb)
        tmp = torch.empty(sh.plan.m_loc, dtype=torch.float64, device=dev)
        L.fill_uniform(tmp, 0x9, offset=sh.plan.r0)
        own.copy_(tmp)
        del tmp
        shards.append(sh)
        xbufs.append(xb)
        ys.append(torch.empty(sh.plan.m_loc, dtype=torch.float64, device=dev))
    torch.cuda.synchronize()

    def step():
        for sh, xb, y in zip(shards, xbufs, ys):
            sh.mul_(y, xb)          # halo exchange (N>1) + lbm_dgbmv on the slab

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream -------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stream = torch.cuda.current_stream()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(BANDS) + 1)] for _ in range(K)]
    launches0 = h.launch_count()
    barrier()
    for k in range(K):
        ev[k][0].record(stream)
        for b, (sh, xb, y) in enumerate(zip(shards, xbufs, ys)):
            sh.mul_(y, xb)
            ev[k][b + 1].record(stream)
    barrier()
    launches = h.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0][0].elapsed_time(ev[K - 1][len(BANDS)])
    per_band_ms = [sum(ev[k][b].elapsed_time(ev[k][b + 1]) for k in range(K)) / K for b in range(len(BANDS))]
    t = torch.tensor([total_ms] + per_band_ms + [float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, per_band_ms = float(tmax[0]), [float(v) for v in tmax[1:1 + len(BANDS)]]
        launches = int(tsum[-1])
    step_bytes = sum(alg_bytes(n, l, u) for (l, u) in BANDS)
    ms_per_step = total_ms / K
    value = step_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (l=u=8 slab apply) ----------------------------------
    peak, peak_src = peaks()
    dom = max(range(len(BANDS)), key=lambda b: per_band_ms[b])
    l, u = BANDS[dom]
    kern_bytes = alg_bytes(n, l, u) / world          # per launch (one launch per rank per apply)
    achieved = kern_bytes / (per_band_ms[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                traffic = json.load(fh).get(f"gbmv_l{l}_n2^{args.log2n or LOG2N}_gpus{world}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": f"gbmv_tile_kernel<double,{l + u + 1}> (l=u={l})",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes,
                "includes_halo_exchange": world > 1}
    breakdown = {f"l{l}u{u}": {"ms": per_band_ms[b], "gbs": alg_bytes(n, l, u) / (per_band_ms[b] * 1e-3) / 1e9,
                                "frac_of_peak": alg_bytes(n, l, u) / (per_band_ms[b] * 1e-3) / 1e9 / (peak * world)}
                 for b, (l, u) in enumerate(BANDS)}

    # ---- the l=u=128 metric point at the largest n that fits (not part of `value`) ------------
    del shards, xbufs, ys
    torch.cuda.empty_cache()
    if not args.no_wide:
        try:
            free = torch.cuda.mem_get_info()[0]
            log2w = min(LOG2N, (int(free * 0.8) * world // (257 * 8)).bit_length() - 1)
            nw = 1 << log2w
            sh = ShardedBanded.brand(nw, 128, 128, rank, world, dtype=torch.float64, device=dev, seed=0x1)
            xb = sh.plan.new_vector(torch.float64, dev)
            L.fill_uniform(xb, 0x9, offset=sh.plan.c0)
            yw = torch.empty(sh.plan.m_loc, dtype=torch.float64, device=dev)
            for _ in range(2):
                sh.mul_(yw, xb)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record(stream)
            for _ in range(reps):
                sh.mul_(yw, xb)
            e1.record(stream)
            barrier()
            msw = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(msw, op=dist.ReduceOp.MAX)
            gbs = alg_bytes(nw, 128, 128) / (float(msw[0]) * 1e-3) / 1e9
            breakdown["l128u128"] = {"n": f"2^{log2w}", "ms": float(msw[0]), "gbs": gbs,
                                     "frac_of_peak": gbs / (peak * world)}
            del sh, xb, yw
            torch.cuda.empty_cache()
        except Exception as e:  # never lose the headline line to the optional point
            breakdown["l128u128"] = {"error": str(e)[:200]}

    # ---- banded x banded materialise, l=u=8 (the other half of the north-star target; not part of `value`):
    #      n = 2^26 f64, columns of C partitioned over the ranks, static halo of A, no communication
    if not args.no_wide:
        ok, err = 1, ""
        try:
            from lbm_b200.shard import GbmmColumnShard
            nm_, lm = 1 << 26, 8
            q = GbmmColumnShard.plan(nm_, lm, lm, lm, lm, rank, world)
            nbm = 2 * lm + 1
            A_slab = torch.empty((q.pb - q.pa, nbm), dtype=torch.float64, device=dev)
            L.fill_uniform(A_slab, 0x1, offset=nbm * q.pa)
            B_slab = torch.empty((q.c1 - q.c0, nbm), dtype=torch.float64, device=dev)
            L.fill_uniform(B_slab, 0x2, offset=nbm * q.c0)
            C_slab = torch.empty((q.c1 - q.c0, 4 * lm + 1), dtype=torch.float64, device=dev)
            for _ in range(2):
                q.local_gbmm(A_slab, B_slab, C_slab)
            torch.cuda.synchronize()
        except Exception as e:  # never lose the headline line to the optional point
            ok, err = 0, str(e)[:200]
        okt = torch.tensor([ok], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)   # every rank takes the same branch (no barrier mismatch)
        if int(okt[0]):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record(stream)
            for _ in range(reps):
                q.local_gbmm(A_slab, B_slab, C_slab)
            e1.record(stream)
            barrier()
            msm = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(msm, op=dist.ReduceOp.MAX)
            gbs = 8 * nm_ * (2 * nbm + 4 * lm + 1) / (float(msm[0]) * 1e-3) / 1e9
            breakdown["materialize_l8u8_x_l8u8"] = {"n": "2^26", "ms": float(msm[0]), "gbs": gbs, "frac_of_peak": gbs / (peak * world),
                                                    "kernel": "gbmm_reg_kernel<double,17,17> (column-sharded, no communication)"}
        else:
            breakdown["materialize_l8u8_x_l8u8"] = {"error": err or "failed on another rank"}
        A_slab = B_slab = C_slab = None
        torch.cuda.empty_cache()

    # ---- e2e: the same metric through the C-ABI host-buffer call (H2D + kernel + D2H timed) ---
    e2e = None if args.no_e2e else run_e2e(args, L, h, n, rank, world, dev, barrier)

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = CpuGbmvSample(24)
        sample.step()
        tb = to = 0.0
        for _ in range(3):
            a_, b_ = sample.step()
            tb += a_
            to += b_
        cpu = {"value": sample.bytes * 3 / tb / 1e9, "unit": "GB/s", "cores": sample.cores, "kind": "port",
               "sample": sample.describe(3), "which": "openblas_dgbmv (the routine the reference's gbmv! ccalls)",
               "oracle_omp_gbs": sample.bytes * 3 / to / 1e9}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"mul!(y,A,x), A=brand(n,n,l,u) f64, n=2^{args.log2n or LOG2N}, l=u in {{1,8}}, "
                                   f"row-sharded over {world} GPU(s), ghost exchange in the timed region",
                       "exchange": "none (1 GPU)" if world == 1 else ("torch.distributed P2P" if args.torch_p2p else
                                                                              ("in-library peer-memory mailboxes (NVLink peer stores + flag handshake kernels), overlapped with interior rows"
                                                                               if peer_on else "in-library ncclSend/ncclRecv, overlapped with interior rows")),
                       "l2": "inputs (3.2 GB and 18.3 GB slabs) far larger than the 126 MB L2; no flush needed",
                       "algorithmic_bytes_per_step": step_bytes},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "breakdown": breakdown, "frac_of_aggregate_peak": value / (peak * world),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(args, L, h, n, rank, world, dev, barrier):
    """Host buffers (pinned) -> lbm_dgbmv_host -> host result, every step: H2D of A and x, the
    kernel, D2H of y.  Each rank pushes its own slab (no halo needed: the host x is global)."""
    import ctypes as C

    import psutil
    import torch
    import torch.distributed as dist
    from lbm_b200.shard import RowShard

    try:
        bands = list(BANDS)
        need = sum((l + u + 1 + 2) * 8 * (n // world + l + u) for (l, u) in bands)
        avail = psutil.virtual_memory().available // max(1, world if world > 1 else 1)
        if need * 1.3 > avail:
            bands = [(1, 1)]
        host = []
        for (l, u) in bands:
            p = RowShard.plan(n, l, u, rank, world)
            lda = l + u + 1
            d = torch.empty((p.n_loc, lda), dtype=torch.float64, device=dev)
            L.fill_uniform(d, 0x1, offset=lda * p.c0)
            xa = torch.empty(p.n_loc, dtype=torch.float64, device=dev)
            L.fill_uniform(xa, 0x9, offset=p.c0)
            hA = torch.empty((p.n_loc, lda), dtype=torch.float64, pin_memory=True)
            hx = torch.empty(p.n_loc, dtype=torch.float64, pin_memory=True)
            hy = torch.empty(p.m_loc, dtype=torch.float64, pin_memory=True)
            hA.copy_(d)
            hx.copy_(xa)
            del d, xa
            host.append((p, hA, hx, hy))
        torch.cuda.empty_cache()
        fn = L.lib().lbm_dgbmv_host

        def step():
            for (p, hA, hx, hy) in host:
                rc = fn(h.ptr, b"N", p.m_loc, p.n_loc, p.kl_loc, p.ku_loc, 1.0, C.c_void_p(hA.data_ptr()),
                        p.l + p.u + 1, C.c_void_p(hx.data_ptr()), 0.0, C.c_void_p(hy.data_ptr()))
                h.check(rc, "lbm_dgbmv_host")

        step()  # warm (allocates the device staging)
        barrier()
        steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(steps):
            step()     # synchronous: returns when y is on the host
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        bytes_alg = sum(alg_bytes(n, l, u) for (l, u) in bands)
        h2d = sum(hA.numel() * 8 + hx.numel() * 8 for (_, hA, hx, _) in host)
        d2h = sum(hy.numel() * 8 for (_, _, _, hy) in host)
        if world > 1:   # whole-job bytes
            tb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
            h2d, d2h = int(tb[0]), int(tb[1])
        checksum = float(sum(hy[:: max(1, hy.numel() // 1024)].sum() for (_, _, _, hy) in host))
        return {"value": bytes_alg / (float(dt[0]) / steps) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": steps, "bands": bands, "result_checksum": checksum,
                "note": "lbm_dgbmv_host with pinned host buffers: H2D(A,x) + kernel + D2H(y) inside the timed region"}
    except Exception as e:
        return {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=0, help="override n = 2^k (testing only; metric is 2^27)")
    ap.add_argument("--no-wide", action="store_true", help="skip the l=u=128 breakdown point")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--nccl-exchange", action="store_true", help="N>1: keep the in-library exchange on ncclSend/ncclRecv "
                    "instead of the peer-memory mailboxes")
    ap.add_argument("--torch-p2p", action="store_true", help="N>1: exchange ghosts through torch.distributed P2P "
                    "instead of the library's own NCCL communicator")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
