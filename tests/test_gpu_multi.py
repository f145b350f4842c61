"""GPU (-m gpu), more than one device: BASELINE config 3 as stated - 16384 x 2^20-bit streams sharded over the GPUs of the
box by the reference's `-j` rule (utilities.c:1526), every batch's p-value records gathered to rank 0 over NCCL - with
the gathered array checked on rank 0 against a single GPU testing all the streams (bit-identical) and against the CPU
oracle on streams sampled from every rank's slice, first and last iteration of every job included
(tools/config3_check.py is the per-rank program).  Skipped on a one-GPU box; run it with `gpurun --gpus N`."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_ranks(nranks, script_args, timeout=1500):
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, "rc %d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-3000:])
    return json.loads(lines[-1])


def test_config3_sharded_gather_matches_single_gpu_and_oracle():
    n = device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    out = run_ranks(world, [os.path.join(ROOT, "tools", "config3_check.py"), "--total", "16384", "--batch", "1024"])
    assert out["ranks"] == world and out["total_streams"] == 16384
    assert out["gathered_identical_to_single_gpu"], out
    assert out["oracle_max_rel_err"] <= 1e-9 and out["ok"], out


def test_ragged_sharding_and_small_batches():
    """a slice that is several short batches per rank (batch 96, 3 x 96 streams per rank): the gather order per batch
    and the reassembly into job order, not only whole 1024-stream batches"""
    n = device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    out = run_ranks(2, [os.path.join(ROOT, "tools", "config3_check.py"), "--total", "576", "--batch", "96", "--oracle-streams", "8"])
    assert out["ok"] and out["total_streams"] == 576, out


def test_sharding_of_a_jump_ahead_generator():
    """the same with generator 2 of tools/generators (512-bit quadratic congruential, stsgpu_generate): every rank jumps to
    its own slice of the one sequence; gathered array against one GPU making all streams, and the oracle"""
    n = device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    out = run_ranks(2, [os.path.join(ROOT, "tools", "config3_check.py"), "--total", "256", "--batch", "64", "--oracle-streams", "6",
                        "--generator", "2"])
    assert out["ok"] and out["generator"] == 2 and out["gathered_identical_to_single_gpu"], out


def test_bench_reports_gather_check_and_strong_scaling():
    """bench.py --gpus N: the gathered p-values of the last timed step are verified on rank 0 (gather_check.ok) and the
    config-3 strong-scaling block is present"""
    n = device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    out = run_ranks(2, [os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4", "--warmup", "3"])
    assert out["n_gpus"] == 2 and out["gather_check"]["ok"] and out["gather_check"]["ranks"] == 2, out.get("gather_check")
    assert out["strong"]["total_streams"] == 16384 and out["strong"]["value"] > 0
