#!/usr/bin/env python
"""bench.py -- banded mul! HBM GB/s & % roofline (BASELINE.json metric).

One "step" = one pass of the hot path over the metric's single-GPU-sized workload:
    mul!(y, A, x) for A = brand(n,n,l,u) Float64, n = 2^27, (l,u) in {(1,1), (8,8)}
row-sharded over the N ranks (one process per GPU; ghost-cell halo exchange of x over NVLink
inside the timed region).  `value` = algorithmic bytes of both applies / step time, aggregated
over all ranks (GB/s).  The (128,128) point of the metric does not fit one GPU at n = 2^27
(A = 276 GB); it is reported in `roofline.configs` (n = 2^27 at N >= 2, else the largest n that fits)
together with every other BASELINE config (C2-C5, materialise), all under one NVML clock sampler.
After the timed loop x is CHANGED and one apply per band is checked on sampled rows -- both ends, random
rows, and the max(l,u) rows either side of every shard boundary -- against the counter-based input
stream (`parity`); a failure exits non-zero.

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
    routine BandedMatrices.gbmv! ccalls -- with every host thread.  `log2n` = 27 is the GPU arm's own config (21.5 GB
    of host RAM); smaller values are a bounded sample.  Inputs are generated once; `step()` times one pass."""

    def __init__(self, log2n):
        import numpy as np
        from oracle import oracle as O
        self.n = 1 << log2n
        self.log2n = log2n
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
        """One pass of OpenBLAS dgbmv over both bands; returns seconds."""
        from scipy.linalg import blas
        tb = 0.0
        for (l, u, A, x, y) in self.ops:
            t0 = time.perf_counter()
            blas.dgbmv(self.n, self.n, l, u, 1.0, A, x, beta=0.0, y=y, overwrite_y=1)
            tb += time.perf_counter() - t0
        return tb

    def step_oracle(self):
        """One pass of the repo's OpenMP oracle port (timed beside the BLAS figure); returns seconds."""
        from oracle import oracle as O
        to = 0.0
        for (l, u, A, x, y) in self.ops:
            t0 = time.perf_counter()
            O.gbmv("N", self.n, self.n, l, u, 1.0, A, x)
            to += time.perf_counter() - t0
        return to

    def describe(self, reps):
        return (f"dgbmv f64 n=2^{self.log2n}, l=u in {{1,8}} ({'the GPU arm' + chr(39) + 's own config' if self.log2n == LOG2N else 'same workload at reduced n'}; "
                f"{reps} passes; OpenBLAS threads={self.blas_threads}, OpenMP threads={self.omp_threads})")


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use all host threads."""
    n = str(os.cpu_count() or 1)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = n


def _host_fits(log2n):
    try:
        import psutil
        need = sum((l + u + 1 + 3) * 8 * (1 << log2n) for (l, u) in BANDS) * 1.25
        return psutil.virtual_memory().available > need
    except Exception:
        return False


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    _use_all_host_threads()          # before numpy / scipy / libgomp are loaded
    log2n = args.log2n or LOG2N
    same = _host_fits(log2n)
    if not same:
        log2n = 24
    sample = CpuGbmvSample(log2n)
    # the n = 2^27 pass costs ~3 s: bound the whole run to a few minutes whatever --steps says
    t_probe = sample.step()
    budget = 150.0
    K = max(1, min(args.steps, int(budget / max(t_probe, 1e-3))))
    W = max(0, min(args.warmup, 2) - 1)
    for _ in range(W):
        sample.step()
    tb = 0.0
    for _ in range(K):
        tb += sample.step()
    to = sample.step_oracle()
    v = sample.bytes * K / tb / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tb / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"mul!(y,A,x), A=brand(n,n,l,u) f64, n=2^{log2n}, l=u in {{1,8}}"
                               + ("" if same else " (bounded sample of the n=2^27 workload: host RAM too small)"),
                   "sharding": "none: one host process, all host threads inside OpenBLAS",
                   "same_config_as_b200_arm": bool(same and log2n == LOG2N), "timed_passes": K,
                   "note": "reference is pure Julia (not runnable here); timed = OpenBLAS dgbmv (scipy.linalg.blas, "
                           "all host threads): the BLAS routine BandedMatrices.gbmv! ccalls, i.e. what the reference "
                           "executes for this path (BASELINE.md section 4). The repo's own OpenMP oracle port is "
                           "timed beside it as cpu_baseline.oracle_omp_gbs."},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": sample.cores, "kind": "port", "sample": sample.describe(K),
                         "which": "openblas_dgbmv", "oracle_omp_gbs": sample.bytes / to / 1e9},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def sample_rows(n, world, halo, k=2000, ends=64, seed=0):
    """Global rows to check: both ends, k random rows, and the halo+2 rows either side of every shard boundary."""
    import numpy as np
    from lbm_b200.shard import split_rows
    parts = [np.arange(0, min(ends, n)), np.arange(max(0, n - ends), n), np.random.default_rng(seed).integers(0, n, k)]
    for r in range(1, world):
        b = split_rows(n, world, r)[0]
        parts.append(np.arange(max(0, b - halo - 2), min(n, b + halo + 2)))
    return np.unique(np.concatenate(parts))


class Parity:
    """Accumulates sampled-entry checks of GPU results against the oracle's counter-based stream restatement."""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world
        self.items = {}
        self.worst = 0.0
        self.checked = 0
        self.ok = True

    def add(self, name, got, want, tol):
        import numpy as np
        import torch
        import torch.distributed as dist
        ratio = float(np.max(np.abs(got - want)) / tol) if len(want) else 0.0
        if len(want) and not np.all(np.isfinite(got)):
            ratio = float("inf")
        t = torch.tensor([ratio, float(len(want))], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            cnt = t[1:].clone()
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            t[1] = cnt[0]
        ratio, cnt = float(t[0]), int(t[1])
        self.items[name] = {"entries": cnt, "max_err_over_bound": ratio, "ok": bool(ratio <= 1.0 and cnt > 0)}
        self.worst = max(self.worst, ratio)
        self.checked += cnt
        self.ok = self.ok and self.items[name]["ok"]

    def report(self):
        return {"ok": bool(self.ok), "max_err_over_bound": self.worst, "entries_checked": self.checked,
                "bound": "4*eps*(l+u+1)*|A|*|x| per row (max-norms), the north-star tolerance",
                "what": "x re-seeded after the timed loop; both ends + random rows + max(l,u)+2 rows either side of every shard boundary, "
                        "recomputed by the oracle from the counter-based input stream", "checks": self.items}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import lbm_b200 as L
    from lbm_b200.shard import (GbmmColumnShard, RowShard, ShardedBanded, ShardedKronSum, ShardedTridiagonal, init_comm,
                                split_batch)
    from oracle import oracle as O          # the checker of the parity leg (never on the timed path)

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
    eps = float(np.finfo(np.float64).eps)
    peak, peak_src = peaks()
    stream = torch.cuda.current_stream()
    par = Parity(dev, world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timeit(fn, reps, warm):
        """max-over-ranks device time per call (CUDA events on the launching stream, barrier on both sides)."""
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def all_ok(ok):
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)   # every rank takes the same branch (no barrier mismatch)
        return bool(int(t[0]))

    def vec(cnt, seed, off=0, dt=torch.float64):
        t = torch.empty(cnt, dtype=dt, device=dev)
        if cnt:
            L.fill_uniform(t, seed, offset=off)
        return t

    def check_banded(name, sh, xb, y, l, u, nn, seed_x):
        """x <- a fresh stream, one sharded apply, sampled rows of this rank's block against the oracle."""
        p = sh.plan
        p.owned(xb).copy_(vec(p.m_loc, seed_x, p.r0))
        y.fill_(float("nan"))
        sh.mul_(y, xb)
        rows = sample_rows(nn, world, max(l, u), seed=l)
        mine = rows[(rows >= p.r0) & (rows < p.r1)]
        got = y[torch.from_numpy(mine - p.r0).to(dev)].cpu().numpy()
        par.add(name, got, O.gbmv_rows_stream(nn, l, u, 0x1, seed_x, mine), 4 * eps * (l + u + 1))

    # ---- resident operands (generated on device from the counter-based stream) -------------
    shards, xbufs, ys = [], [], []
    for (l, u) in BANDS:
        sh = ShardedBanded.brand(n, l, u, rank, world, dtype=torch.float64, device=dev, seed=0x1)
        xb = sh.plan.new_vector(torch.float64, dev)
        sh.plan.owned(xb).copy_(vec(sh.plan.m_loc, 0x9, sh.plan.r0))
        shards.append(sh)
        xbufs.append(xb)
        ys.append(torch.empty(sh.plan.m_loc, dtype=torch.float64, device=dev))
    torch.cuda.synchronize()

    def step():
        for sh, xb, y in zip(shards, xbufs, ys):
            sh.mul_(y, xb)          # halo exchange (N>1) + lbm_dgbmv on the slab

    for _ in range(W):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream -------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
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

    # ---- parity of the timed path: x changed, boundary rows included ---------------------------
    if not args.no_parity:
        for (l, u), sh, xb, y in zip(BANDS, shards, xbufs, ys):
            check_banded(f"M gbmv l=u={l} n=2^{args.log2n or LOG2N}", sh, xb, y, l, u, n, 0x51 + l)

    # ---- roofline of the dominant kernel (l=u=8 slab apply) ----------------------------------
    dom = max(range(len(BANDS)), key=lambda b: per_band_ms[b])
    l, u = BANDS[dom]
    kern_bytes = alg_bytes(n, l, u) / world          # per launch (one launch per rank per apply)
    achieved = kern_bytes / (per_band_ms[dom] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            with open(tpath) as fh:
                traffic = json.load(fh).get(f"gbmv_l{l}_n2^{args.log2n or LOG2N}_gpus{world}")
                traffic_src = "constant from the committed ncu --set full capture (profiles/ncu_M_r02.md), not measured in this run"
        except Exception:
            traffic = None
    configs = []

    def cfg(name, ms, byts=None, flops=None, **kw):
        r = {"config": name, "ms": ms}
        if byts is not None:
            r["value"] = byts / ms / 1e6
            r["unit"] = "GB/s"
            r["frac"] = r["value"] / (peak * world)
        if flops is not None:
            r["tflops"] = flops / ms / 1e9
        r.update(kw)
        configs.append(r)
        return r

    for b, (bl, bu) in enumerate(BANDS):
        cfg(f"M gbmv f64 n=2^{args.log2n or LOG2N} l=u={bl} (timed region)", per_band_ms[b], alg_bytes(n, bl, bu),
            launches_per_apply=launches / (world * K * len(BANDS)))

    del shards, xbufs, ys
    torch.cuda.empty_cache()

    # ---- every other BASELINE config, all under ONE clock sampler ------------------------------
    sampler2 = ClockSampler(local)
    if rank == 0 and not args.no_wide:
        sampler2.start()

    def guarded(name, fn):
        """run one config; a failure on any rank is recorded, never loses the headline line"""
        ok, err = True, ""
        try:
            fn()
        except Exception as e:   # noqa: BLE001
            ok, err = False, f"{type(e).__name__}: {e}"[:200]
        if not all_ok(ok):
            configs.append({"config": name, "error": err or "failed on another rank"})
        torch.cuda.empty_cache()

    def c_wide():
        # the l=u=128 point of the metric: n = 2^27 where it fits (N >= 2), else the largest power of two that does
        free = torch.cuda.mem_get_info()[0]
        log2w = min(LOG2N, (int(free * 0.8) * world // (257 * 8)).bit_length() - 1)
        nw = 1 << log2w
        sh = ShardedBanded.brand(nw, 128, 128, rank, world, dtype=torch.float64, device=dev, seed=0x1)
        xb = sh.plan.new_vector(torch.float64, dev)
        sh.plan.owned(xb).copy_(vec(sh.plan.m_loc, 0x9, sh.plan.r0))
        yw = torch.empty(sh.plan.m_loc, dtype=torch.float64, device=dev)
        ms = timeit(lambda: sh.mul_(yw, xb), 3, 2)
        cfg(f"M gbmv f64 n=2^{log2w} l=u=128" + ("" if log2w == LOG2N else " (n=2^27 needs >= 2 GPUs: A = 276 GB)"), ms,
            alg_bytes(nw, 128, 128), kernel="gbmv_wide_n_kernel<double>")
        if not args.no_parity:
            check_banded(f"M gbmv l=u=128 n=2^{log2w}", sh, xb, yw, 128, 128, nw, 0x5F)

    def c_gm():
        # banded x banded materialise, l=u=8 (the other half of the north-star target): n = 2^26 f64, columns of C
        # partitioned over the ranks, static halo of A, no communication
        nm_, lm = 1 << 26, 8
        q = GbmmColumnShard.plan(nm_, lm, lm, lm, lm, rank, world)
        nbm = 2 * lm + 1
        A_slab = torch.empty((q.pb - q.pa, nbm), dtype=torch.float64, device=dev)
        L.fill_uniform(A_slab, 0x1, offset=nbm * q.pa)
        B_slab = torch.empty((q.c1 - q.c0, nbm), dtype=torch.float64, device=dev)
        L.fill_uniform(B_slab, 0x2, offset=nbm * q.c0)
        C_slab = torch.empty((q.c1 - q.c0, 4 * lm + 1), dtype=torch.float64, device=dev)
        ms = timeit(lambda: q.local_gbmm(A_slab, B_slab, C_slab), 5, 2)
        cfg("GM gbmm f64 n=2^26 (8,8)x(8,8)->(16,16), column-sharded, no communication", ms, 8 * nm_ * (2 * nbm + 4 * lm + 1),
            2 * nm_ * nbm * nbm, kernel="gbmm_reg_kernel<double,17,17>")
        if not args.no_parity:
            _check_gbmm_window(f"GM gbmm (8,8)^2 n=2^26", q, C_slab, nm_, lm, nbm)

    def _check_gbmm_window(name, q, C_slab, nn, lm, nbm, ncols=32):
        """a window of `ncols` columns at the start of this rank's block (i.e. at the shard boundary) recomputed by the oracle"""
        c0 = q.c0
        w = GbmmColumnShard(nn, lm, lm, lm, lm, 0, 1, c0, min(nn, c0 + ncols))
        (kla, kua), (klb, kub), (klc, kuc) = w.sub_bandwidths()
        As = np.asfortranarray(O.fill_uniform(1, nbm * (w.pb - w.pa), offset=nbm * w.pa).reshape((nbm, w.pb - w.pa), order="F"))
        Bs = np.asfortranarray(O.fill_uniform(2, nbm * (w.c1 - w.c0), offset=nbm * c0).reshape((nbm, w.c1 - w.c0), order="F"))
        want = O.gbmm(w.rb - w.ra, w.c1 - w.c0, w.pb - w.pa, As, kla, kua, Bs, klb, kub, klc=klc, kuc=kuc)
        got = np.asfortranarray(C_slab[: w.c1 - w.c0].cpu().numpy().T)
        m_loc = w.rb - w.ra
        dg, dw = O.band_to_dense(got, m_loc, klc, kuc), O.band_to_dense(want, m_loc, klc, kuc)
        par.add(name, dg.ravel(), dw.ravel(), 4 * eps * nbm)

    def c2():
        # C2: SymTridiagonal / Bidiagonal mat-vec n = 2^28, row-sharded with ghost exchange
        n2 = 1 << 28
        for nm, has_l, has_u, sl, su in (("SymTridiagonal", True, True, 2, 2), ("Bidiagonal U", False, True, None, 2)):
            c0, c1 = ShardedTridiagonal.slices(n2, int(has_l), int(has_u), rank, world)
            d = vec(c1 - c0, 1, c0)
            e = vec(c1 - c0 - 1, 2, c0)
            sh = ShardedTridiagonal(n2, e if has_l else None, d, e if has_u else None, rank, world)
            p = sh.plan
            xb = p.new_vector(torch.float64, dev)
            p.owned(xb).copy_(vec(p.m_loc, 4, p.r0))
            y = torch.empty(p.m_loc, dtype=torch.float64, device=dev)
            l0 = h.launch_count()
            ms = timeit(lambda: sh.mul_(y, xb), 10, 3)
            cfg(f"C2 {nm} mul! f64 n=2^28, row-sharded", ms, 4 * 8 * n2, kernel="tridiag_ring_kernel<double>",
                launches_per_apply=(h.launch_count() - l0) / 13)
            if not args.no_parity:
                p.owned(xb).copy_(vec(p.m_loc, 0x44, p.r0))
                y.fill_(float("nan"))
                sh.mul_(y, xb)
                rows = sample_rows(n2, world, 1, seed=3)
                mine = rows[(rows >= p.r0) & (rows < p.r1)]
                got = y[torch.from_numpy(mine - p.r0).to(dev)].cpu().numpy()
                par.add(f"C2 {nm} n=2^28", got, O.tridiag_rows_stream(n2, sl, 1, su, 0x44, mine), 4 * eps * 3)
            del d, e, sh, xb, y
            torch.cuda.empty_cache()

    def c3():
        # C3: 2-D Laplacian on 8192^2, partitioned by grid columns, never materialised
        g = 8192
        c2_ = float(g) ** 2
        D = L.banded_from_diagonals(g, {0: -2.0 * c2_, 1: c2_, -1: c2_}, device=dev)
        p = RowShard.plan(g, 1, 1, rank, world)
        sh = ShardedKronSum(g, g, 1, 1, D.data[p.c0:p.c1].contiguous(), D, rank, world)
        xb = sh.new_buffer(torch.float64, dev)
        L.fill_uniform(sh.owned(xb).view(-1), 5, offset=p.r0 * g)
        y = torch.empty((p.m_loc, g), dtype=torch.float64, device=dev)
        l0 = h.launch_count()
        # 200 applies per timed window (8 ms at N=8): at 40 us per apply a 20-apply window still sits on the ramp of the first
        # launches (benchmarks/c3_multi_probe.py: 34.7 us per apply over 20, 30.9 us over 200, same build, N=2)
        ms = timeit(lambda: sh.mul_(y, xb), 200, 20)
        cfg("C3 2-D Laplacian kron(D,I)+kron(I,D) 8192x8192, grid-column-sharded", ms, 16 * g * g, kernel="kron_ring_kernel<double,3,3>",
            launches_per_apply=(h.launch_count() - l0) / 220, applies_timed=200)
        if not args.no_parity:
            L.fill_uniform(sh.owned(xb).view(-1), 0x55, offset=p.r0 * g)
            y.fill_(float("nan"))
            sh.mul_(y, xb)
            cols = sample_rows(g, world, 1, k=48, ends=2, seed=5)
            cols = cols[(cols >= p.r0) & (cols < p.r1)]
            rr = np.unique(np.concatenate([np.arange(0, 4), np.arange(g - 4, g), np.random.default_rng(6).integers(0, g, 40)]))
            C_, R_ = np.meshgrid(cols, rr, indexing="ij")
            got = y[torch.from_numpy(C_.ravel() - p.r0).to(dev), torch.from_numpy(R_.ravel()).to(dev)].cpu().numpy()
            par.add("C3 Laplacian 8192^2", got, O.laplacian_points_stream(g, c2_, 0x55, R_.ravel(), C_.ravel()),
                    4 * eps * 6 * 2 * c2_)

    def c4():
        # C4: gbmm l=u=128, n = 2^20, column-partitioned (static halo of A, no communication): FP64 DMMA tiles
        n4 = 1 << 20
        q = GbmmColumnShard.plan(n4, 128, 128, 128, 128, rank, world)
        A_slab = torch.empty((q.pb - q.pa, 257), dtype=torch.float64, device=dev)
        L.fill_uniform(A_slab, 1, offset=257 * q.pa)
        B_slab = torch.empty((q.c1 - q.c0, 257), dtype=torch.float64, device=dev)
        L.fill_uniform(B_slab, 2, offset=257 * q.c0)
        C_slab = torch.empty((q.c1 - q.c0, 513), dtype=torch.float64, device=dev)
        ms = timeit(lambda: q.local_gbmm(A_slab, B_slab, C_slab), 3, 2)
        flops = 2 * n4 * 257 * 257
        # FP64 denominators: the DMMA issue-loop ceiling measured by benchmarks/micro/fp64_pipes.cu (37.1 TFLOP/s per GPU) and
        # cuBLAS DGEMM 8192^3 timed here, in this run (library call used only as a denominator)
        a_ = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        c_ = torch.empty_like(a_)
        for _ in range(2):
            torch.matmul(a_, a_, out=c_)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            torch.matmul(a_, a_, out=c_)
        e1.record(stream)
        torch.cuda.synchronize()
        dgemm = 3 * 2 * 8192**3 / e0.elapsed_time(e1) / 1e9
        tf = flops / ms / 1e9
        cfg("C4 gbmm f64 n=2^20 (128,128)x(128,128)->(256,256), column-sharded, no communication", ms, None, flops,
            bound="tensor", value=tf, unit="TFLOP/s", peak_dmma_loop=37.1 * world, frac=tf / (37.1 * world),
            cublas_dgemm_8192_tflops_this_run=dgemm, frac_of_cublas_dgemm=tf / (dgemm * world),
            hbm_gbs=8 * n4 * (257 + 257 + 513) / ms / 1e6, kernel="gbmm_dmma_ws_kernel (mma.sync DMMA.8x8x4)")
        del a_, c_
        if not args.no_parity:
            _check_gbmm_window("C4 gbmm (128,128)^2 n=2^20", q, C_slab, n4, 128, 257, ncols=16)

    def c5():
        # C5: 4096 fused chains A*B*C, brand(4096,4096,4,4) f32, batch split, no communication
        batch, nn = 4096, 4096
        b0, b1 = split_batch(batch, world, rank)
        fs = [torch.empty((b1 - b0, nn, 9), dtype=torch.float32, device=dev) for _ in range(3)]
        for i, f in enumerate(fs):
            L.fill_uniform(f, 1 + i, offset=b0 * nn * 9)
        out = [None]

        def run():
            out[0] = L.chain3_batched(*fs, [4, 4, 4], [4, 4, 4])
        ms = timeit(run, 5, 2)
        cfg("C5 fused chain A*B*C f32 batch=4096 n=4096 (4,4)^3, batch-split", ms, batch * 4 * nn * (9 + 9 + 9 + 25),
            batch * 2 * nn * (9 * 9 + 17 * 9), kernel="gbmm_chain3_warp_kernel<float>")
        if not args.no_parity and b1 > b0:
            items = [O.fill_uniform(1 + i, nn * 9, np.float32, offset=b0 * nn * 9).reshape(1, nn, 9) for i in range(3)]
            want = O.gbmm_chain3_batched(nn, [4, 4, 4], [4, 4, 4], *items)[0]
            got = out[0][0].cpu().numpy()
            e32 = float(np.finfo(np.float32).eps)
            # stage bounds add: 9 terms of O(1) then 9 terms of O(9) (tests/test_gpu_fullsize.py)
            par.add("C5 chain item at the batch-split boundary", got.ravel().astype(np.float64), want.ravel().astype(np.float64),
                    4 * e32 * 9 * 9 + 4 * e32 * 9 * 9)

    if not args.no_wide:
        # M-128 goes LAST: it holds up to 0.8 of the device memory (138 GB per GPU at N=2), and whatever is timed right after
        # that much memory is released read 76 % instead of 91 % in some processes (GM at N=2, rounds 1 and 2; the same
        # per-rank problem alone on a GPU is a steady 91 %, benchmarks/gm_n2_probe.py)
        for nm, fn in (("GM", c_gm), ("C2", c2), ("C3", c3), ("C4", c4), ("C5", c5), ("M-128", c_wide)):
            guarded(nm, fn)
    clocks2 = sampler2.stop() if (rank == 0 and not args.no_wide) else None

    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": f"gbmv_tile_kernel<double,{l + u + 1}> (l=u={l})",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes,
                "includes_halo_exchange": world > 1, "configs": configs, "configs_clocks": clocks2,
                "configs_note": "per config: max-over-ranks device ms per apply (exchange included), aggregate algorithmic GB/s, "
                                "frac = share of n_gpus x the measured per-GPU peak; C4: TFLOP/s against n_gpus x 37.1 (DMMA loop) "
                                "and against cuBLAS DGEMM timed in this run"}

    # ---- e2e: the same metric through the C-ABI host-buffer call (H2D + kernel + D2H timed) ---
    e2e = None if args.no_e2e else run_e2e(args, L, h, n, rank, world, dev, barrier)

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = CpuGbmvSample(24)
        sample.step()
        tb = sum(sample.step() for _ in range(3))
        to = sample.step_oracle()
        cpu = {"value": sample.bytes * 3 / tb / 1e9, "unit": "GB/s", "cores": sample.cores, "kind": "port",
               "sample": sample.describe(3), "which": "openblas_dgbmv (the routine the reference's gbmv! ccalls)",
               "oracle_omp_gbs": sample.bytes / to / 1e9}

    parity = None if args.no_parity else par.report()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"mul!(y,A,x), A=brand(n,n,l,u) f64, n=2^{args.log2n or LOG2N}, l=u in {{1,8}}",
                       "sharding": f"row-sharded over {world} GPU(s), ghost exchange in the timed region",
                       "exchange": "none (1 GPU)" if world == 1 else ("torch.distributed P2P" if args.torch_p2p else
                                                                              ("in-library peer-memory mailboxes (NVLink peer stores + flag handshake) "
                                                                               if peer_on else "in-library ncclSend/ncclRecv, overlapped with interior rows")),
                       "l2": "inputs (3.2 GB and 18.3 GB slabs) far larger than the 126 MB L2; no flush needed",
                       "l128": "the metric's l=u=128 band is not part of `value` (it does not fit one GPU at n=2^27, and `value` must mean "
                               "the same at every N); it is roofline.configs['M gbmv ... l=u=128']",
                       "algorithmic_bytes_per_step": step_bytes},
            "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "frac_of_aggregate_peak": value / (peak * world),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        if rank == 0:
            print("PARITY FAILURE: " + json.dumps(parity["checks"]), file=sys.stderr)
        return 3
    return 0


def _bind_numa_local(local):
    """Pin this process to the host cores of the NUMA node its GPU hangs off BEFORE the pinned buffers are allocated and
    first touched, so every rank's H2D stream reads node-local DRAM (torchrun does not bind ranks).  Returns a note."""
    try:
        import torch
        bdf = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            hd = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            bdf = pynvml.nvmlDeviceGetPciInfo(hd).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        except Exception:
            return "numa: no NVML"
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node_path = f"/sys/bus/pci/devices/{bdf}/numa_node"
        if not os.path.exists(node_path):
            return f"numa: {node_path} missing"
        node = int(open(node_path).read().strip())
        if node < 0:
            return "numa: node unknown (-1)"
        cl = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return f"numa: node {node} cpus not in this cgroup"
        os.sched_setaffinity(0, use)
        return f"numa: bound to node {node} ({len(use)} cpus)"
    except Exception as e:   # noqa: BLE001
        return f"numa: {type(e).__name__}"


def run_e2e(args, L, h, n, rank, world, dev, barrier):
    """Host buffers (pinned) -> lbm_dgbmv_host -> host result, every step: H2D of A and x, the
    kernel, D2H of y.  Each rank pushes its own slab (no halo needed: the host x is global).
    `resident` = the drop-in use north_star describes: A uploaded once, per step H2D(x) + kernel + D2H(y)."""
    import ctypes as C

    import psutil
    import torch
    import torch.distributed as dist
    from lbm_b200.shard import RowShard

    try:
        numa_note = _bind_numa_local(dev.index) if not args.no_numa else "numa: binding off"
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
            host.append((p, hA, hx, hy, d, xa))
        fn = L.lib().lbm_dgbmv_host

        def step():
            for (p, hA, hx, hy, _, _) in host:
                rc = fn(h.ptr, b"N", p.m_loc, p.n_loc, p.kl_loc, p.ku_loc, 1.0, C.c_void_p(hA.data_ptr()),
                        p.l + p.u + 1, C.c_void_p(hx.data_ptr()), 0.0, C.c_void_p(hy.data_ptr()))
                h.check(rc, "lbm_dgbmv_host")

        def timed(fn_step, steps):
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn_step()     # synchronous: returns when y is on the host
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt[0]) / steps

        step()  # warm (allocates the device staging)
        steps = max(1, min(args.steps, 3))
        dt = timed(step, steps)
        bytes_alg = sum(alg_bytes(n, l, u) for (l, u) in bands)
        h2d = sum(hA.numel() * 8 + hx.numel() * 8 for (_, hA, hx, _, _, _) in host)
        d2h = sum(hy.numel() * 8 for (_, _, _, hy, _, _) in host)
        h2d_local = h2d

        # resident operator: A stays in HBM; per step H2D(x) + kernel + D2H(y)
        gfn = L.lib().lbm_dgbmv
        lib = L.lib()

        def step_res():
            for (p, hA, hx, hy, d, xa) in host:
                h.check(lib.lbm_memcpy_h2d(h.ptr, C.c_void_p(xa.data_ptr()), C.c_void_p(hx.data_ptr()), hx.numel() * 8), "h2d")
                h.check(gfn(h.ptr, b"N", p.m_loc, p.n_loc, p.kl_loc, p.ku_loc, 1.0, C.c_void_p(d.data_ptr()), p.l + p.u + 1,
                            C.c_void_p(xa.data_ptr()), 0.0, C.c_void_p(ybufs[id(p)].data_ptr())), "lbm_dgbmv")
                h.check(lib.lbm_memcpy_d2h(h.ptr, C.c_void_p(hy.data_ptr()), C.c_void_p(ybufs[id(p)].data_ptr()), hy.numel() * 8), "d2h")
            h.sync()

        ybufs = {id(p): torch.empty(p.m_loc, dtype=torch.float64, device=dev) for (p, *_r) in host}
        h.use_torch_stream()
        step_res()
        dt_res = timed(step_res, max(3, steps))
        h2d_res = sum(hx.numel() * 8 for (_, _, hx, _, _, _) in host)

        # the H2D stream alone, per rank (where the N >= 2 plateau of round 1 comes from)
        def step_copy():
            for (p, hA, hx, hy, d, xa) in host:
                h.check(lib.lbm_memcpy_h2d(h.ptr, C.c_void_p(d.data_ptr()), C.c_void_p(hA.data_ptr()), hA.numel() * 8), "h2d")
            h.sync()
        step_copy()
        dt_copy = timed(step_copy, 2)
        copy_gbs = sum(hA.numel() * 8 for (_, hA, *_r) in host) / dt_copy / 1e9

        if world > 1:   # whole-job bytes
            tb = torch.tensor([float(h2d), float(d2h), float(h2d_res)], dtype=torch.float64, device=dev)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
            h2d, d2h, h2d_res = int(tb[0]), int(tb[1]), int(tb[2])
        checksum = float(sum(hy[:: max(1, hy.numel() // 1024)].sum() for (_, _, _, hy, _, _) in host))
        return {"value": bytes_alg / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": steps, "bands": bands, "result_checksum": checksum,
                "note": "lbm_dgbmv_host with pinned host buffers: H2D(A,x) + kernel + D2H(y) inside the timed region",
                "h2d_gbs_per_rank_alone": copy_gbs, "h2d_gbs_per_rank_in_e2e": h2d_local / dt / 1e9, "numa": numa_note,
                "resident": {"value": bytes_alg / dt_res / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d_res, "d2h_bytes_per_step": d2h,
                             "ms_per_step": dt_res * 1e3,
                             "note": "operator resident in HBM (uploaded once, the drop-in use north_star describes); per step "
                                     "lbm_memcpy_h2d(x) + lbm_dgbmv + lbm_memcpy_d2h(y) + lbm_sync, pinned host buffers"}}
    except Exception as e:   # noqa: BLE001
        return {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": f"{type(e).__name__}: {e}"[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=0, help="override n = 2^k (testing only; metric is 2^27)")
    ap.add_argument("--no-wide", action="store_true", help="skip roofline.configs (l=u=128, GM, C2-C5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing parity leg (profiling runs)")
    ap.add_argument("--no-numa", action="store_true", help="e2e: do not bind the rank to its GPU's NUMA node")
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
