#!/usr/bin/env python3
"""bench.py - Gbit/s tested, all 15 SP800-22 tests on 2^20-bit streams (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (all 15 test kernels + p-value tails) over one batch of
1024 synthetic bitstreams of 2^20 bits per GPU (BASELINE config 2; 128 MiB of packed input per step,
larger than L2).  Streams are the reference's own LCG (tools/generators 1), generated on the device by
skip-ahead so that rank r owns streams [r*1024, (r+1)*1024) - the reference's `-j` sharding rule.

  value   whole-job Gbit/s with the batch already resident in HBM: kernels + tails + D2H of the
          records (+ NCCL gather of the p-values to rank 0 when N > 1), max over ranks; two batches in flight.
  e2e     the same metric through the C ABI with HOST input: stsgpu_submit() from pinned host
          memory (H2D inside), wait, records on the host, gather; three batches in flight.
  roofline  dominant kernel group of the step (the DFT): algorithmic bytes / its CUDA-event duration with the
          groups serialised, taken in a short pass right after the timed region on the same inputs; the same
          events inside the timed region, where batches and streams overlap, are reported as in_region_ms.
  cpu_baseline  the unmodified reference binary (oracle/_ref/sts, built from /root/reference with the
          fftw3 shim) on all host cores, on a bounded sample of the same workload (rank 0, N = 1).

  gather_check  (N > 1) rank 0 compares the gathered p-values of the last step, rank by rank in `-j` order, with what
          it computes itself for those streams (bit-identical) and with the CPU oracle on a few sampled streams.
  strong  BASELINE config 3 next to the weak-scaling headline: 16384 x 2^20-bit streams IN TOTAL sharded over the
          N GPUs by the `-j` rule, p-values gathered to rank 0 (strong scaling: fixed total work).
  roofline_all  every kernel group against its own roofline class.

`--impl reference` times that reference binary alone (all host cores, bounded sample per step).
The product path never touches oracle/: it is only executed here as the reported baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BITS = 1 << 20
STREAMS_PER_GPU = 1024
REF_STREAMS_PER_CORE = 8
FFT_NOTE = ("libfftw3 is not installed in this image: the reference binary is linked against the plain FP64 FFT of "
            "oracle/fft_f64.c through a 5-symbol fftw3.h shim, so its DFT leg (about 2 % of its time) is slower than with FFTW")
CONFIG3_TOTAL = 16384
METRIC = "Gbit/s tested, all 15 tests on 2^20-bit streams"
WORKLOAD = "1024x2^20-bit LCG streams per GPU, all 15 tests, default parameters (BASELINE config 2)"


def csrc_sha256():
    """hash of the engine's CUDA sources: ncu figures under profiles/ are only quoted for the build they were taken on"""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "sts_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "sts_b200", "csrc", "*.cuh"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                mhz.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def barrier(dist):
    if dist is not None:
        import torch
        torch.cuda.synchronize()
        dist.barrier()


def max_over_ranks(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "sts")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def lcg_file(path, streams):
    """write `streams` LCG streams as the reference's generator would (oracle restatement, pinned to
    tools/generators by tests/golden/make_golden.py)"""
    from oracle import oracle as O
    O.lcg_bytes(0, streams).tofile(path)


def time_reference(streams, cores, workdir):
    """wall seconds of `sts -v 1 -i streams -F r -T cores` on an LCG file (BASELINE config 1 shape)"""
    data = os.path.join(workdir, "lcg_%d.bin" % streams)
    if not os.path.exists(data):
        lcg_file(data, streams)
    ref = reference_binary()
    if ref is not None:
        w = tempfile.mkdtemp(prefix="stsref_", dir=workdir)
        cmd = [ref, "-v", "1", "-i", str(streams), "-F", "r", "-T", str(cores), "-w", w, data]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        if r.returncode == 0:
            return dt, "reference"
    # reference binary not shipped: time the C restatement instead (same algorithm, same threading model)
    import numpy as np
    from oracle import oracle as O
    raw = np.fromfile(data, dtype=np.uint8)
    t0 = time.perf_counter()
    O.run(O.make_params(), raw, streams, nthreads=cores)
    return time.perf_counter() - t0, "port"


def run_reference_arm(args, rank, world, dist):
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = REF_STREAMS_PER_CORE * cores			# several streams per thread: the wall time is not one slow thread's
    tmp = tempfile.mkdtemp(prefix="stsbench_")
    times = []
    kind = "reference"
    for i in range(args.warmup_ref + args.steps_ref):
        dt, kind = time_reference(sample, cores, tmp)
        if i >= args.warmup_ref:
            times.append(dt)
    sec = sum(times) / len(times)
    gbit = sample * N_BITS / sec / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbit, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d streams per step" % sample},
        "cpu_baseline": {"value": gbit, "unit": "Gbit/s", "cores": cores, "kind": kind,
                         "sample": "%d x 2^20-bit LCG streams per step, sts -i %d -F r -T %d (file read included); %s"
                                   % (sample, sample, cores, FFT_NOTE)},
        "e2e": {"value": gbit, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_config4(sts_b200, device, streams=428, steps=4):
    """BASELINE config 4 as a measured configuration: `streams` LCG streams of 10^7 bits (4 x the 107 streams that
    1024 Mi-bit of generator output hold), -P 6=5000,2=10,4=16,5=16; BlockFrequency self-disables at this n
    (blockFrequency.c:119), 323 p-values per stream.  Resident input, two batches in flight; per-group times from a
    serialised pass afterwards."""
    n = 10 ** 7
    eng = sts_b200.Engine(device=device, max_streams=streams, n=n, linear_complexity_M=5000, non_overlapping_m=10,
                          apen_m=16, serial_m=16, pipeline_depth=2)
    try:
        for d in range(2):
            eng.generate_lcg(0, streams)
            eng.submit_resident(streams, 0)
        eng.wait()
        eng.wait()
        for i in range(2):                                   # warm-up
            eng.submit_resident(streams, 0)
            eng.wait()
        t0 = time.perf_counter()
        pending = 0
        for i in range(steps):
            eng.submit_resident(streams, 0)
            pending += 1
            if pending == 2:
                eng.wait()
                pending -= 1
        while pending:
            eng.wait()
            pending -= 1
        sec = (time.perf_counter() - t0) / steps
        eng.set_profiling(1)
        acc = {}
        for i in range(3):
            eng.submit_resident(streams, 0)
            eng.wait()
            if i >= 1:
                for name, ms in eng.kernel_times():
                    acc.setdefault(name, []).append(ms)
        return {"workload": "%d x 10^7-bit LCG streams, -P 6=5000,2=10,4=16,5=16 (BASELINE config 4), 14 tests, %d p-values per stream"
                            % (streams, eng.layout.npv),
                "streams_per_step": streams, "bits_per_stream": n, "steps": steps, "ms_per_step": sec * 1e3,
                "ms_per_107_streams": sec * 1e3 * 107.0 / streams, "value": streams * n / sec / 1e9, "unit": "Gbit/s",
                "kernel_groups_ms_serialised": {k: round(sum(v) / len(v), 3) for k, v in acc.items()}}
    finally:
        eng.close()


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the BASELINE config-4 block")
    args = ap.parse_args()

    rank, world, local, dist = (0, 1, 0, None)
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        # bounded: a handful of multi-second runs, independent of the GPU arm's K/W
        args.steps_ref = max(1, min(args.steps, 3))
        args.warmup_ref = 1 if args.warmup > 0 else 0
        return run_reference_arm(args, rank, world, dist)

    rank, world, local, dist = dist_setup(args.gpus)
    import numpy as np
    import sts_b200

    B = args.streams
    K, W = args.steps, max(args.warmup, 3)
    DEPTH = int(os.environ.get("STS_BENCH_DEPTH", "3"))		# batches in flight (stsgpu_params.pipeline_depth)
    eng = sts_b200.Engine(device=local, max_streams=B, pipeline_depth=DEPTH)	# raises if libstsgpu.so or the GPU is missing
    lay = eng.layout
    if world > 1:
        import torch
        uid = [sts_b200.Engine.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
        # gather mode: the exchange of every batch's p-values to rank 0 is queued behind its tail kernels and
        # overlaps the next batches; stsgpu_wait returns with the gathered array on rank 0's host
        eng.gather_begin(0)

    # synthetic input: rank r owns LCG streams [r*B, (r+1)*B)  (the reference's -j partition),
    # generated into every batch slot so that the resident runs start with the input in HBM
    pinned = []
    for d in range(DEPTH):
        eng.generate_lcg(rank * B, B)
        pb = sts_b200.binding.PinnedBuffer(B * N_BITS // 8)
        pb.array[:] = eng.download_input(B)
        pinned.append(pb)
        eng.submit_resident(B, rank * B)
    for d in range(DEPTH):
        eng.wait()

    # assessment tallies on the device (stsgpu_assess_*): every step below also adds its 1024 x 188 p-values
    # to the per-column uniformity / proportion tallies; they are summed over ranks and judged after the timing
    BINS = 32
    eng.assess_begin(BINS)

    last_gather = [None]

    def finish_one(sink=None):
        eng.wait()
        if sink is not None:
            sink()
        if world > 1:
            last_gather[0] = eng.gather_pvals(0, world, copy=False)	# rank 0: all ranks' records of this step (pinned view)

    def run_steps(count, submit, sink=None, inflight=None):
        """exactly `count` batches submitted and completed, at most `inflight` (default DEPTH) in flight"""
        pending = 0
        inflight = inflight or DEPTH
        for i in range(count):
            submit(i)
            pending += 1
            if pending == inflight:
                finish_one(sink)
                pending -= 1
        while pending:
            finish_one(sink)
            pending -= 1

    def submit_resident(i):
        eng.submit_resident(B, rank * B)

    def submit_host(i):
        eng.submit(pinned[i % DEPTH].ptr, B, rank * B)

    # ---- value: resident input ----
    # two batches in flight: with the input already in HBM a third batch only adds contention (measured 9.35 ms per
    # batch against 9.5); the host-input loop below uses all DEPTH slots so that one batch's upload is always hidden
    RES_INFLIGHT = min(DEPTH, 2)
    run_steps(W, submit_resident, None, RES_INFLIGHT)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ktimes_region = {}
    dev_ms = []

    def sink_times():
        dev_ms.append(eng.last_batch_ms())
        for name, ms in eng.kernel_times():
            ktimes_region.setdefault(name, []).append(ms)

    barrier(dist)
    t0 = time.perf_counter()
    run_steps(K, submit_resident, sink_times, RES_INFLIGHT)
    barrier(dist)
    t1 = time.perf_counter()
    launches = eng.launch_count() - launches0
    sec = max_over_ranks(dist, t1 - t0)
    dev_step_ms = max_over_ranks(dist, sum(dev_ms) / len(dev_ms))
    value = world * B * N_BITS * K / sec / 1e9

    # ---- e2e: host input through the C ABI (H2D of every batch and D2H of its records inside) ----
    def sink_read():
        pv, st, w = eng.results(copy=False)
        sink_read.acc += float(pv[0, 0])			# touch the records on the host
    sink_read.acc = 0.0
    run_steps(2, submit_host, sink_read)
    barrier(dist)
    t0 = time.perf_counter()
    run_steps(K, submit_host, sink_read)
    barrier(dist)
    t1 = time.perf_counter()
    sec_e2e = max_over_ranks(dist, t1 - t0)
    clocks = sampler.stop() if rank == 0 else None
    gathered_last = last_gather[0].copy() if (world > 1 and rank == 0) else None
    if world > 1:
        eng.gather_begin(-1)					# what follows is local to every rank
    e2e = world * B * N_BITS * K / sec_e2e / 1e9
    h2d = B * N_BITS // 8
    d2h = B * (lay.npv * 8 + lay.nst * 8 + lay.nw * 4)

    # ---- assessment: one all-reduce of 188 x 34 counters, then the verdicts (igamc on the device) ----
    eng.assess_reduce()
    tallied = int(eng.assess_finish()["sample"][lay.pv_off[1]])	# every step above was tallied (the same streams again and again)
    eng.assess_begin(BINS)					# verdicts over distinct streams: one batch per rank, sqrt(1024) = 32 bins
    eng.submit_resident(B, rank * B)
    eng.wait()
    eng.assess_reduce()
    verdict = eng.assess_finish()
    assessment = {"uniformity_bins": BINS, "columns": int(lay.npv), "streams_tallied_in_timed_steps": tallied,
                  "streams_assessed": int(verdict["sample"][lay.pv_off[1]]),
                  "columns_passing_uniformity_and_proportion": int((verdict["flags"] == 3).sum())}

    eng.assess_begin(BINS)					# the steps below tally again, like the timed ones

    # ---- gather_check (N > 1): the gathered records of the last timed step, rank by rank in -j order ----
    gather_check = None
    if world > 1:
        if rank == 0:
            same, worst = 0, 0.0
            assert gathered_last.shape == (world * B, lay.npv), gathered_last.shape
            for r in range(world):				# rank r tested streams [r*B, (r+1)*B): recompute them here
                eng.generate_lcg(r * B, B)
                eng.submit_resident(B, r * B)
                eng.wait()
                pv_r = eng.results(copy=False)[0]
                same += int((gathered_last[r * B:(r + 1) * B] == pv_r).all())
            # the CPU oracle as the checker (outside every timed region): one stream of every rank's block
            from oracle import oracle as O
            idx = [r * B + (37 * r + 11) % B for r in range(world)]
            raws = []
            for i in idx:			# the device generator (pinned to tools/generators by the tests) has skip-ahead, the CPU one has not
                eng.generate_lcg(i, 1)
                raws.append(eng.download_input(1))
            raw = np.concatenate(raws)
            opv = O.run(O.make_params(), raw, len(idx), nthreads=os.cpu_count() or 1)[0]
            got = gathered_last[idx]
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = np.abs(got - opv) / np.abs(opv)
            rel[got == opv] = 0.0
            c3 = slice(lay.pv_off[3], lay.pv_off[3] + 2)
            rel[:, c3][np.abs(got[:, c3] - opv[:, c3]) <= 1e-13] = 0.0	# cusum: 1 - sum + sum (tests/test_gpu_parity.py)
            worst = float(rel.max())
            gather_check = {"ranks": world, "streams": world * B, "blocks_identical_to_rank0_recompute": same,
                            "oracle_streams": idx, "oracle_max_rel_err": worst,
                            "ok": bool(same == world and worst <= 1e-9)}
        barrier(dist)

    # ---- strong scaling: BASELINE config 3, 16384 streams in total over the N GPUs, gathered to rank 0 ----
    strong = None
    if (CONFIG3_TOTAL // world) % B == 0:
        from sts_b200.sharding import run_sharded
        run_sharded(eng, world * B * 2, rank, world, B, collect=False)		# warm-up: two batches per rank
        marks = []
        barrier(dist)
        run_sharded(eng, CONFIG3_TOTAL, rank, world, B, collect=False, timer=lambda: marks.append(time.perf_counter()))
        barrier(dist)
        sec3 = max_over_ranks(dist, marks[1] - marks[0])
        strong = {"workload": "%d x 2^20-bit LCG streams in total (2 GiB), sharded over %d GPU(s) by the -j rule, all 15 tests, "
                              "p-values gathered to rank 0 (BASELINE config 3)" % (CONFIG3_TOTAL, world),
                  "scaling": "strong", "total_streams": CONFIG3_TOTAL, "n_gpus": world, "seconds": sec3,
                  "value": CONFIG3_TOTAL * N_BITS / sec3 / 1e9, "unit": "Gbit/s",
                  "input": "generated on the device batch by batch (generator included in the time)"}

    # ---- per-group durations with every group alone on the GPU (CUDA events, serialised pass) ----
    ktimes = {}
    eng.set_profiling(1)
    for i in range(5):
        eng.submit_resident(B, rank * B)
        eng.wait()
        if i >= 2:
            for name, ms in eng.kernel_times():
                ktimes.setdefault(name, []).append(ms)
    eng.set_profiling(0)

    # ---- roofline of the dominant kernel group (CUDA events on its launch stream, timed region) ----
    kavg = {k: sum(v) / len(v) for k, v in ktimes.items()}
    top = max(kavg, key=kavg.get)
    hbm_peak, peak_src = peaks()
    stream_bytes = N_BITS // 8
    # algorithmic bytes per stream of every HBM-class group: the packed stream once per kernel;
    # the DFT adds its FP64 spectrum written and read once; Universal reads only (Q + K) L bits
    hbm_bytes = {"dft": 2 * (N_BITS // 2) * 16 + stream_bytes,
                 "scan(freq,cusum,runs,excursions)": stream_bytes,
                 "blocks(blockfreq,longestrun,overlapping)": 3 * stream_bytes,
                 "universal": 113120}
    alg_bytes = B * hbm_bytes.get(top, stream_bytes)
    achieved = alg_bytes / (kavg[top] * 1e-3) / 1e9
    # what ncu measured (DRAM bytes, pipe utilisation per kernel) is only quoted while the kernels are the ones that
    # were profiled: profiles/r2_ncu_summary.json records a hash of sts_b200/csrc at capture time
    ncu = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")))
        if tj.get("source_sha256") == csrc_sha256() and tj.get("bits") == N_BITS:
            ncu = tj
    except Exception:
        pass
    traffic = None
    if ncu and top in ncu.get("dram_bytes_per_group", {}):
        traffic = ncu["dram_bytes_per_group"][top] * B / ncu["streams"]
    roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": kavg[top], "algorithmic_bytes_per_launch_group": alg_bytes,
                "all_kernel_groups_ms": {k: round(v, 4) for k, v in kavg.items()},
                "in_region_ms": {k: round(sum(v) / len(v), 4) for k, v in ktimes_region.items()},
                "note": "kernel_ms / all_kernel_groups_ms: CUDA events around each kernel group with the groups "
                        "serialised (3 steps right after the timed region, same inputs); in_region_ms: the same "
                        "events inside the timed region, where several batches and five streams per batch overlap; "
                        "traffic: DRAM bytes from the ncu capture under profiles/, null when the kernels changed since"}
    roofline_all = []
    for name, ms in sorted(kavg.items(), key=lambda kv: -kv[1]):
        e = {"kernel": name, "ms": round(ms, 4)}
        if name in hbm_bytes:
            a = B * hbm_bytes[name] / (ms * 1e-3) / 1e9
            e.update({"bound": "hbm", "achieved": round(a, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(a / hbm_peak, 4)})
        elif name.startswith("tails"):
            e.update({"bound": "latency (igamc / erfc dependency chains, 1.5 KB of records per stream)", "achieved": None,
                      "peak": None, "unit": None, "frac": None})
        else:
            # integer ALU (LinearComplexity, Rank) or shared-memory atomics (pattern histograms, templates): the
            # utilisation of the limiting pipe as ncu measured it, quoted only for the profiled build
            pipe = (ncu or {}).get("pipe_util_pct", {}).get(name)
            e.update({"bound": "int-pipe" if name in ("linear_complexity", "rank") else "smem-atomic (LSU) pipe",
                      "achieved": pipe, "peak": 100.0 if pipe is not None else None, "unit": "% of pipe (ncu)" if pipe is not None else None,
                      "frac": round(pipe / 100.0, 4) if pipe is not None else None})
        roofline_all.append(e)

    # ---- BASELINE config 4 (n = 10^7, LinearComplexity M = 5000, template m = 10, Serial/ApEn m = 16) on this GPU ----
    config4 = None
    if rank == 0 and not args.no_config4:
        config4 = run_config4(sts_b200, local)
    barrier(dist)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = 4 * cores
        tmp = tempfile.mkdtemp(prefix="stsbench_")
        dt, kind = time_reference(sample, cores, tmp)
        cpu = {"value": sample * N_BITS / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": kind,
               "sample": "%d x 2^20-bit LCG streams, `sts -i %d -F r -T %d` wall time %.1f s (file read included); %s"
                         % (sample, sample, cores, dt, FFT_NOTE)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": sec / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams_per_gpu_per_step": B, "bits_per_stream": N_BITS,
                       "l2": "input per step (128 MiB) exceeds L2; the FFT intermediate (8 GiB per step) flushes it",
                       "pipeline": "%d batch slots per GPU: 2 in flight for resident input, %d for host input (uploads run that far ahead, the kernels of at most two batches interleave)" % (DEPTH, DEPTH),
                       "sharding": "rank r tests streams [r*B,(r+1)*B), p-values gathered to rank 0 over NCCL in gather mode "
                                   "(the exchange of a batch is queued behind its kernels and overlaps the next batches)"},
            "device_ms_per_step": dev_step_ms,
            "e2e": {"value": e2e, "unit": "Gbit/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": sec_e2e / K * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_all": roofline_all,
            "gather_check": gather_check,
            "strong": strong,
            "config4": config4,
            "cpu_baseline": cpu,
            "assessment": assessment,
        }
        print(json.dumps(line))
    for pb in pinned:
        pb.free()
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
