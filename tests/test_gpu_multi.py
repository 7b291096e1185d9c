"""Multi-GPU parity on hardware: spawns one rank per GPU (torchrun, NCCL) when the box shows >= 2 GPUs, skipped
otherwise.  Covers what a single-GPU `pytest -m gpu` cannot: the in-library ghost exchange (peer mailboxes fused into
the compute launch, the unfused two-kernel exchange, and ncclSend/ncclRecv), sharded chains, the Kronecker-sum column
partition, short / empty trailing shards, and bench.py's own post-timing parity leg at N = 2."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(nproc, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), script, *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_parity_on_gpus(nproc):
    if _ngpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    r = _torchrun(nproc, os.path.join("benchmarks", "check_shard_nccl.py"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK sharded" in r.stdout


def test_bench_parity_leg_two_gpus():
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", "--log2n", "24", "--no-e2e", "--no-cpu", "--no-wide")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["parity"]["ok"] and line["parity"]["entries_checked"] > 4000
    assert line["n_gpus"] == 2


def test_single_process_multi_device():
    """one host thread driving every GPU of the box through lbm_multi_* (no torch.distributed)."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join("benchmarks", "check_multi_single_process.py")], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK single-process" in r.stdout
