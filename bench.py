#!/usr/bin/env python
"""
bench.py -- headline benchmark of the Lightworks emulator probability hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the largest configuration that fits one GPU; SURVEY.md 8d "C3"):
    Sampler / Permanent backend, 24-mode Haar-random unitary (scipy unitary_group, seed 3), input
    |1^10, 0^14>, ALL 92 561 040 output-state probabilities |Perm(U_{S,T})|^2 / (prod s! prod t!).
One step = one pass over that distribution.  With N GPUs the output Fock states are sharded by
contiguous rank ranges of the reference's fock_basis order (strong scaling: total work fixed); the
step then ends with the NCCL all-gather of the probability shards and the scalar all-reduce of their
sum (the normalisation check), as BASELINE.json's north_star describes.

One JSON line on stdout (rank 0).  Keys follow the driver contract; see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M, N_PH, SEED = 24, 10, 3
WORKLOAD = ("C3: Sampler/Permanent backend, 24-mode Haar unitary (seed 3), 10 photons |1^10,0^14>, "
            "all 92,561,040 output-state probabilities")
METRIC = "output-state probabilities/sec (complex128 10x10 permanents, full distribution)"
UNIT = "probabilities/s"


def average_terms(m: int, n: int) -> float:
    """Average number of terms of the multiplicity-aware walk over ALL outputs of n photons in m modes:
    prod(s_i + 1) / (2 if some s_i is odd else 1) per multiplicity pattern, weighted by its number of states."""
    from math import comb, factorial

    def parts(k, mx):
        if k == 0:
            yield []
            return
        for p in range(min(k, mx), 0, -1):
            for rest in parts(k - p, p):
                yield [p, *rest]

    tot = work = 0
    for part in parts(n, n):
        d = len(part)
        if d > m:
            continue
        cnt = factorial(m) // factorial(m - d)
        for v in {x: part.count(x) for x in part}.values():
            cnt //= factorial(v)
        t = 1
        for s_ in part:
            t *= s_ + 1
        if any(s_ % 2 for s_ in part):
            t //= 2
        tot += cnt
        work += cnt * t
    assert tot == comb(m + n - 1, n)
    return work / tot


def flops_per_perm(n: int) -> float:
    """Algorithmic FP64 flops of one n x n complex128 Glynn/Gray permanent."""
    return 2.0 ** (n - 1) * (8 * n - 4)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float | None = None, t1: float | None = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = [r for (ts, r) in self.rows if t0 is None or (t0 <= ts <= t1 + 0.05)]
        scope = "timed region"
        if not window:  # region shorter than the sampling period: use every sample taken under load (warm-up + region)
            window = [r for (_, r) in self.rows[1:]] or [r for (_, r) in self.rows]
            scope = "warm-up + timed region (region shorter than the sampling period)"
        for r in window:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_port_baseline(U, ins, sample: int, threads: int) -> dict:  # noqa: N803
    """The oracle's C restatement of the reference loop (permanent.py:145-150), all host threads, on a
    bounded prefix of the same workload."""
    from oracle import oracle as O  # noqa: N812  (bench.py's cpu_baseline leg may execute oracle/)

    O.basis_probabilities(U, ins, 0, 64, threads=threads)  # warm
    t0 = time.perf_counter()
    O.basis_probabilities(U, ins, 0, sample, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample} of 92561040 output states in fock_basis order, oracle/lw_oracle.c "
                      f"(restated thewalrus perm_bbfg), {threads} pthreads, {dt:.2f} s"}


def run_reference_arm(args) -> int:
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O  # noqa: N812
    from oracle import refshim

    U = O.random_unitary(M, SEED)  # noqa: N806
    ins = [1] * N_PH + [0] * (M - N_PH)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic", "config": {"workload": WORKLOAD, "modes": M, "photons": N_PH}}
    times = []
    if refshim.reference_available():
        # the unmodified reference (pure Python, single-threaded) through its own public method:
        # PermanentBackend.probability per output state, exactly the loop body of permanent.py:145-150
        refshim.install()
        from lightworks.emulator.backends.permanent import PermanentBackend as RefPermanent  # noqa: PLC0415

        sample = args.ref_sample or 20000
        outs = O.fock_basis(M, N_PH, limit=sample).tolist()  # same order as fock_basis(24,10)[:sample]
        backend = RefPermanent()
        for step in range(args.warmup + args.steps):
            n_do = min(sample, 2000) if step < args.warmup else sample
            t0 = time.perf_counter()
            for o in outs[:n_do]:
                backend.probability(U, ins, o)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
        kind, cores = "reference", 1
        desc = (f"unmodified lightworks {refshim.REFERENCE_ROOT} PermanentBackend.probability over the first {sample} "
                "of 92561040 outputs per step (thewalrus.perm restated in C, see oracle/refshim); single thread: "
                "the reference has no parallelism")
    else:
        threads = os.cpu_count() or 1
        sample = args.ref_sample or 200000
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.basis_probabilities(U, ins, 0, sample if step >= args.warmup else 2000, threads=threads)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
        kind, cores = "port", threads
        desc = f"oracle/lw_oracle.c port over the first {sample} outputs per step, {threads} pthreads"
    total = sum(times)
    value = sample * len(times) / total
    line.update({"value": value, "ms_per_step": 1e3 * total / len(times),
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def protect_stdout():
    """The driver parses ONE JSON line from stdout; NCCL ("NCCL version ...") and other libraries print there
    too.  Route fd 1 to stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def haar_unitary(n: int, seed: int):
    """scipy.stats.unitary_group.rvs(n, random_state=seed): what lightworks.random_unitary(n, seed) returns
    (sdk/utils/random.py:66-67) - the synthetic input of every workload here."""
    from scipy.stats import unitary_group

    return unitary_group.rvs(n, random_state=seed)


class _Inputs:  # the only thing the timed legs need from an input generator
    random_unitary = staticmethod(haar_unitary)


def main() -> int:  # noqa: PLR0915
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=400000)
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import lightworks_b200 as lb
    from lightworks_b200 import _lib
    O = _Inputs  # noqa: N806  Haar inputs; oracle/ is only imported by the cpu_baseline and --impl reference legs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)

    ctx = lb.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    U = O.random_unitary(M, SEED)  # noqa: N806  == lw.random_unitary(24, seed=3) (sdk/utils/random.py:66-67)
    ins = np.array([1] * N_PH + [0] * (M - N_PH), dtype=np.uint8)
    S = lb.fock_count(M, N_PH)  # noqa: N806
    # N > 1: the PRODUCT's multi-GPU entry point (include/lwb200.h "multi-GPU", lightworks_b200.Comm): contiguous
    # rank-range shards of equal WORK (lwb_shard_cuts), every probability stored straight into the peers' copies of the
    # distribution by the permanent kernel (NVLink peer stores, CUDA IPC between the ranks), one all-reduce of the sum
    # that doubles as the completion barrier.  Every rank ends each step with the WHOLE distribution, contiguous, in
    # fock_basis order.  bench.py only broadcasts the NCCL unique id and times the call.
    comm = None
    if world > 1:
        uid = [lb.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = lb.Comm.rank(ctx, rank, world, uid[0])
        if os.environ.get("LWB_BENCH_PEER_STORES"):
            comm.set_option("peer_stores", int(os.environ["LWB_BENCH_PEER_STORES"]))
        r0, r1 = comm.shard(M, N_PH, rank)
    else:
        r0, r1 = 0, S
    cnt = r1 - r0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        d_U = torch.from_numpy(U).cuda()  # noqa: N806  resident inputs for the device-timed region
        d_ins = torch.from_numpy(ins).cuda()
        d_probs = torch.zeros(cnt, dtype=torch.float64, device="cuda")  # this rank's shard (N = 1: everything)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    comm_out = {}

    def device_step():
        if world == 1:
            ctx.perm_basis_probs_dev(d_U, M, d_ins, N_PH, r0, cnt, d_probs)
        else:
            comm_out["probs"], comm_out["sum"] = comm.perm_basis_probs_dev(d_U, M, d_ins, N_PH)

    # pinned host buffers for the end-to-end leg
    h_U = torch.from_numpy(U.copy()).pin_memory()  # noqa: N806
    h_ins = torch.from_numpy(ins.copy()).pin_memory()
    h_out = torch.empty(S if world > 1 else cnt, dtype=torch.float64).pin_memory()
    h_U_np, h_ins_np, h_out_np = h_U.numpy(), h_ins.numpy(), h_out.numpy()  # noqa: N806

    def e2e_step():
        # the public host-buffer API: H2D of U + input, kernels (N > 1: incl. the peer-store exchange and the
        # all-reduce, so every rank's DEVICE copy is complete), D2H of this rank's shard into its slice of the host
        # array, sync.  (N > 1: one process per GPU - each rank's host holds its own rank range.)
        if world == 1:
            ctx.perm_basis_probs(h_U_np, h_ins_np, r0, cnt, out=h_out_np)
        else:
            comm.perm_basis_probs(h_U_np, h_ins_np, gather_host=False, out=h_out_np)

    def timed(step_fn, k, sampler=None):
        """K steps between barrier+synchronize; per-step CUDA events on the launch stream; the L2 flush
        between steps is not inside any event pair.  Returns (sum of event ms, wall ms)."""
        evs = []
        barrier()
        t_start = time.time()
        w0 = time.perf_counter()
        with torch.cuda.stream(stream):
            for _ in range(k):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_fn()
                e1.record(stream)
                evs.append((e0, e1))
        barrier()
        wall = (time.perf_counter() - w0) * 1e3
        clocks = sampler.stop(t_start, time.time()) if sampler else None
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return ms, wall, clocks

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- warm-up, then the device-resident timed region ----
    sampler = ClockSampler(local)
    sampler.start()  # started before the warm-up so that samples exist even when the timed region is tens of ms
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            device_step()
    l0 = ctx.launch_count()
    dev_ms, dev_wall, clocks = timed(device_step, args.steps, sampler)
    launches = ctx.launch_count() - l0
    per_rank_ms = None
    if world > 1:
        # balance diagnostic: every rank's own shard kernels WITHOUT peer stores or collectives (3 repetitions, best)
        best = 1e30
        with torch.cuda.stream(stream):
            ctx.perm_basis_probs_dev(d_U, M, d_ins, N_PH, r0, cnt, d_probs)  # warm: sizes the scratch for this shard
            stream.synchronize()
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.perm_basis_probs_dev(d_U, M, d_ins, N_PH, r0, cnt, d_probs)
                e1.record(stream)
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
        t_all = torch.zeros(world, dtype=torch.float64, device="cuda")
        t_all[rank] = best
        dist.all_reduce(t_all)
        per_rank_ms = [round(float(v), 3) for v in t_all]
    dev_ms = max_over_ranks(dev_ms)
    if world == 1:
        total_sum = float(d_probs.sum())
    else:
        # every rank holds the whole distribution, contiguous: its own shard bit-identical to the stand-alone launch,
        # the sum of the complete copy equal to the all-reduced sum of the shard sums
        full = torch.as_tensor(comm_out["probs"], device="cuda")
        total_sum = float(torch.as_tensor(comm_out["sum"], device="cuda")[0])
        assert full.numel() == S
        assert torch.equal(full[r0:r1], d_probs), "sharded result differs from the stand-alone shard"
        assert abs(float(full.sum()) - total_sum) < 1e-9, (float(full.sum()), total_sum)
        assert abs(total_sum - 1.0) < 1e-9, total_sum
    value = S * args.steps / (dev_ms * 1e-3)

    # ---- end-to-end through the host-buffer C ABI ----
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = S * args.steps / e2e_s

    # ---- roofline of the dominant kernel ----
    # The default path walks the multiplicity-aware Glynn sum (K1x: prod(s_i+1)/2 terms for an output with
    # occupations s_i); the plain expanded-row kernel (K1: 2^(n-1) terms for every output) is timed beside it.
    # `achieved` is ALGORITHMIC flops (2^(n-1)(8n-4) per permanent) over the kernel time for both;
    # `executed_frac` counts the FP64 work K1x really issues (terms x (8n-2)).
    peak = ctx.fp64_peak()  # measured live: MEASURED_PEAKS.json carries no FP64 entry

    def kernel_ms(reps=3):
        with torch.cuda.stream(stream):
            ctx.perm_basis_probs_dev(d_U, M, d_ins, N_PH, r0, cnt, d_probs)  # warm-up: scratch sized, caches primed
            stream.synchronize()
            kev = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.perm_basis_probs_dev(d_U, M, d_ins, N_PH, r0, cnt, d_probs)
                e1.record(stream)
                kev.append((e0, e1))
            stream.synchronize()
        return statistics.mean(a.elapsed_time(b) for a, b in kev)

    k_ms = kernel_ms()
    lb.set_option("multiplicity_aware", 0)
    try:
        with torch.cuda.stream(stream):
            ctx.perm_basis_probs_dev(d_U, M, d_ins, N_PH, r0, cnt, d_probs)
        plain_ms = kernel_ms()
    finally:
        lb.set_option("multiplicity_aware", 1)
    alg = cnt * flops_per_perm(N_PH)
    avg_terms = average_terms(M, N_PH)
    executed = cnt * avg_terms * (8 * N_PH - 2)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    peak_src = ("lwb_fp64_peak DFMA micro-benchmark run inside this bench (MEASURED_PEAKS.json has no FP64 entry; nominal "
                "148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s)")
    roofline = {"bound": "fp64", "kernel": "k1x_fused_kernel_2<N=10> + bucketing passes (multiplicity-aware Glynn/Gray walk, one launch for all patterns)",
                "achieved": alg / (k_ms * 1e-3) / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                "frac": alg / (k_ms * 1e-3) / peak, "executed_frac": executed / (k_ms * 1e-3) / peak,
                "average_terms_per_permanent": avg_terms, "terms_expanded": 2 ** (N_PH - 1),
                "traffic": traffic,
                "traffic_note": "DRAM bytes of one fused-kernel launch (ncu): 4x the 740 MB of algorithmic output because the "
                                "probabilities leave in bucket (multiplicity-pattern) order as scattered 8-byte stores "
                                "(partial-sector fills) and the 370 MB bucket index is read; 3.0 GB per 96 ms = 31 GB/s, "
                                "0.5 % of the HBM peak - the kernel is FP64-bound",
                "kernel_ms": k_ms, "flops_per_unit": flops_per_perm(N_PH), "units_per_launch": cnt,
                "peak_source": peak_src,
                "note": "frac uses the reference algorithm's flop count per permanent, so a walk that skips the repeated-row "
                        "terms of bunched outputs can exceed the 0.57 register-file ceiling of the expanded-row kernel; "
                        "tensor cores unused (not a dense contraction)"}
    roofline_plain = {"bound": "fp64", "kernel": "perm_states_kernel<N=10> (expanded rows, 2^(n-1) Gray terms per output)",
                      "achieved": alg / (plain_ms * 1e-3) / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                      "frac": alg / (plain_ms * 1e-3) / peak, "kernel_ms": plain_ms,
                      "note": "ceilings: (8n-4)/(12n-4) = 0.655 if every FP64 instruction issued in 2 cycles; 0.57 measured "
                              "limit because a DFMA with three distinct register operands issues in 3 (lwb_fp64_probe)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "modes": M, "photons": N_PH, "outputs": S,
                   "parallelism": f"fock-rank-range shards x{world} (equal work)" + (
                       f", lwb_comm_perm_basis_probs_dev: {'in-kernel NVLink peer stores into every rank' if comm.info()['peer_stores'] else 'in-place NCCL broadcasts of the shards'}"
                       " + NCCL all-reduce of the sum; every rank ends with the whole contiguous distribution" if world > 1 else ""),
                   "l2": "256 MiB flush between timed steps (outside the event pairs); outputs 740 MB > 126 MB L2",
                   "check_sum_of_probabilities": total_sum},
        "roofline": roofline, "roofline_expanded_rows": roofline_plain,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(U.nbytes + ins.nbytes),
                "d2h_bytes_per_step": int(cnt * 8), "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": ("lwb_perm_basis_probs (host buffers, pinned) via lightworks_b200.Context.perm_basis_probs" if world == 1 else
                        "lwb_comm_perm_basis_probs (host buffers, pinned; device copies complete on every rank, each rank "
                        "copies its own rank range to its host) via lightworks_b200.Comm.perm_basis_probs")},
        "gpu_launches": int(launches), "clocks": clocks, "wall_ms_timed_region": dev_wall,
    }
    if per_rank_ms is not None:
        line["per_rank_compute_ms"] = per_rank_ms  # shard kernels only, no collectives

    if world > 1 and not args.no_extras:
        # C5: single n x n permanents of the 60-mode Haar unitary (seed 5), Gray-code space sliced over the ranks by
        # lwb_comm_perm_single_dev (16-byte NCCL all-reduce inside the call); the same permanent on ONE GPU beside it
        U60 = O.random_unitary(60, 5)  # noqa: N806
        c5 = {}
        for n5 in (20, 22, 24, 26, 28):
            with torch.cuda.stream(stream):
                A5 = torch.from_numpy(U60[:n5, :n5].copy()).cuda()  # noqa: N806
                o5 = torch.zeros(2, dtype=torch.float64, device="cuda")
                o1 = torch.zeros(2, dtype=torch.float64, device="cuda")

            def single_step():
                comm.perm_single_dev(A5, n5, o5)

            def single_one_gpu():
                ctx.perm_single_dev(A5, n5, o1)

            with torch.cuda.stream(stream):
                single_step()
                single_one_gpu()
            ms5, _, _ = timed(single_step, 5)
            ms5 = max_over_ranks(ms5) / 5
            ms1, _, _ = timed(single_one_gpu, 5)
            ms1 = max_over_ranks(ms1) / 5
            assert abs(complex(float(o5[0]), float(o5[1])) - complex(float(o1[0]), float(o1[1]))) <= 1e-10 * abs(complex(float(o1[0]), float(o1[1])))
            c5[f"n{n5}"] = {"ms": ms5, "ms_one_gpu": ms1, "speedup": ms1 / ms5, "efficiency": ms1 / ms5 / world,
                            "permanents_per_s": 1e3 / ms5,
                            "fp64_frac_per_gpu": flops_per_perm(n5) / (ms5 * 1e-3) / peak / world,
                            "path": "Gray slices + 16-byte all-reduce" if n5 >= 24 else
                                    "every rank walks the whole space (below 24 photons the exchange costs more than it saves)"}
        line["extra"] = {"single_perm_sharded": c5}
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        line["cpu_baseline"] = cpu_port_baseline(U, ins, args.cpu_sample, threads)
        if not args.no_extras:
            try:
                line["extra"] = extras(ctx, stream, peak, d_U, torch, np, O, _lib)
            except Exception as e:  # the extras never take the headline line down with them
                line["extra"] = {"error": repr(e)}
            sl = line["extra"].get("slos_c3") if isinstance(line["extra"], dict) else None
            if sl and "probabilities_per_s" in sl:
                # second headline: the SAME distribution through the SLOS backend (Backend("slos")), HBM-bound
                line["second_headline"] = {
                    "metric": "output-state probabilities/sec (SLOS backend: complex128 state-vector evolution, full distribution)",
                    "value": sl["probabilities_per_s"], "unit": UNIT, "ms_per_step": sl["ms"], "roofline": sl["roofline"],
                    "kernel": "slos_units4_kernel (K3r: prefix units, coalesced parent streams, parent block in shared memory, "
                              "host-built records instead of index arithmetic) x 10 layers",
                    "gather_kernel_ms": sl.get("gather_kernel_ms"), "unit_kernel_ms": sl.get("unit_kernel_ms")}
    if rank == 0:
        emit(line)
    if world > 1:
        barrier()
        comm.close()  # collective: every rank unmaps its peers before anybody frees
        dist.destroy_process_group()
    return 0


def extras(ctx, stream, peak, d_U, torch, np, O, _lib) -> dict:  # noqa: N803
    """Other rows of the metric (reported, not the headline): SLOS probabilities/s on the same distribution
    with its HBM roofline, batched n x n permanents/s (complex128 n = 12, 16, 20; complex64 n = 12), the C4
    noise-model hot path (all 255 photon-subset inputs of 8 photons in 16 modes) and single large permanents."""
    out = {}
    hbm = None
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0  # B200_PROFILING.md fallback

    def t_ms(fn, reps=3):
        with torch.cuda.stream(stream):
            fn()
            stream.synchronize()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
        return min(ts)

    from math import comb

    S = _lib.fock_count(M, N_PH)  # noqa: N806
    ins = [1] * N_PH + [0] * (M - N_PH)
    with torch.cuda.stream(stream):
        d_p = torch.empty(S, dtype=torch.float64, device="cuda")
    ms = t_ms(lambda: ctx.slos_basis_probs_dev(d_U, M, ins, d_p, order=1), reps=5)
    ms_gather = ms_units = None
    try:  # beside it: the round-1 gather-form kernel and the index-deriving unit kernel (same results up to rounding)
        _lib.set_option("slos_units", 0)
        ms_gather = t_ms(lambda: ctx.slos_basis_probs_dev(d_U, M, ins, d_p, order=1))
        _lib.set_option("slos_units", 1)
        ms_units = t_ms(lambda: ctx.slos_basis_probs_dev(d_U, M, ins, d_p, order=1))
    finally:
        _lib.set_option("slos_units", 4)
    # algorithmic bytes: every layer reads its parent once and writes itself once (the last layer
    # writes 8-byte probabilities)
    slos_bytes = sum(16 * (comb(M + k - 2, k - 1) + comb(M + k - 1, k)) for k in range(1, N_PH + 1))
    slos_traffic = None
    try:  # DRAM bytes of the whole evolution from the committed ncu launch list
        slos_traffic = json.load(open(os.path.join(ROOT, "profiles", "k3_traffic.json"))).get("dram_bytes_per_evolution")
    except Exception:
        slos_traffic = None
    out["slos_c3"] = {"probabilities_per_s": S / (ms * 1e-3), "ms": ms, "gather_kernel_ms": ms_gather,
                      "unit_kernel_ms": ms_units,
                      "roofline": {"bound": "hbm", "achieved": slos_bytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                   "frac": slos_bytes / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": slos_bytes,
                                   "traffic": slos_traffic},
                      "gather_traffic": {"bytes": 16 * M * sum(comb(M + k - 2, k - 1) for k in range(1, N_PH + 1)),
                                         "achieved_gbs": 16 * M * sum(comb(M + k - 2, k - 1) for k in range(1, N_PH + 1)) / (ms * 1e-3) / 1e9,
                                         "note": "every (child, occupied mode) edge of a layer reads a distinct parent "
                                                 "amplitude: m*S(m,k-1) 16-byte gathers per layer through L1/L2; the DRAM "
                                                 "figure above assumes each parent is fetched once"},
                      "note": "SLOS backend, same distribution in SLOS dict order, all 10 layers"}
    # on-device sampling of the same distribution: CDF build + search, device-resident, and the
    # whole Sampler chain through the host API (U in, 10^6 variates in, 10^6 drawn ranks out)
    n_smp = 1_000_000
    with torch.cuda.stream(stream):
        ctx.perm_basis_probs_dev(d_U, M, torch.tensor(ins, dtype=torch.uint8, device="cuda"), N_PH, 0, S, d_p)
        d_u = torch.rand(n_smp, dtype=torch.float64, device="cuda")
        d_i = torch.empty(n_smp, dtype=torch.int64, device="cuda")
    ms = t_ms(lambda: ctx.sample_dev(d_p, S, d_u, n_smp, d_i, threshold=1e-9))
    scan_bytes = 24 * S  # two reads of the probabilities, one write of the CDF
    u_host = np.random.default_rng(0).random(n_smp)
    U_host = O.random_unitary(M, 3)  # noqa: N806
    ctx.basis_sample(U_host, ins, u_host)
    e2e_ms = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.basis_sample(U_host, ins, u_host)
        e2e_ms = min(e2e_ms, 1e3 * (time.perf_counter() - t0))
    out["sampler_c3_on_device"] = {"samples": n_smp, "cdf_and_search_ms": ms,
                                   "roofline": {"bound": "hbm", "achieved": scan_bytes / (ms * 1e-3) / 1e9, "peak": hbm,
                                                "unit": "GB/s", "frac": scan_bytes / (ms * 1e-3) / 1e9 / hbm,
                                                "algorithmic_bytes": scan_bytes,
                                                "note": "CDF of 92.6 M probabilities (24 B per state) + 10^6 binary searches"},
                                   "e2e_ms": e2e_ms, "e2e_samples_per_s": n_smp / (e2e_ms * 1e-3),
                                   "note": "lwb_basis_sample: distribution computed, thresholded, normalised and sampled on "
                                           "the GPU; host traffic = U + 8 B per variate in, 8 B per sample out"}
    del d_p, d_u, d_i
    for n, dt, name in ((12, _lib.LWB_C128, "c128"), (16, _lib.LWB_C128, "c128"), (20, _lib.LWB_C128, "c128"),
                        (12, _lib.LWB_C64, "c64")):
        B = int(max(2e11 / flops_per_perm(n), 148 * 64))  # noqa: N806
        with torch.cuda.stream(stream):
            A = torch.randn(B, n, n, 2, dtype=torch.float64, device="cuda") / np.sqrt(n)  # noqa: N806
            o = torch.empty(B, 2, dtype=torch.float64, device="cuda")
        ms = t_ms(lambda: ctx.perm_matrices_dev(A, B, n, o, dt))
        rec = {"permanents_per_s": B / (ms * 1e-3), "batch": B, "ms": ms}
        if dt == _lib.LWB_C128:
            rec["fp64_frac"] = B * flops_per_perm(n) / (ms * 1e-3) / peak
        out[f"batched_perm_n{n}_{name}"] = rec
        del A, o
    # C4 hot path: every non-empty subset of the 8 input photons (255 inputs), 16 modes, full output basis each
    U16 = torch.from_numpy(O.random_unitary(16, 4)).cuda()  # noqa: N806
    subsets = [[(b >> i) & 1 for i in range(8)] + [0] * 8 for b in range(1, 256)]
    with torch.cuda.stream(stream):
        d_sub = [torch.tensor(sb, dtype=torch.uint8, device="cuda") for sb in subsets]
        d_buf = torch.empty(_lib.fock_count(16, 8), dtype=torch.float64, device="cuda")
    counts = [_lib.fock_count(16, sum(sb)) for sb in subsets]

    def c4():
        for sb, dsb, c in zip(subsets, d_sub, counts):
            ctx.perm_basis_probs_dev(U16, 16, dsb, sum(sb), 0, c, d_buf)

    ms = t_ms(c4)
    out["c4_hot_path"] = {"permanents_per_s": sum(counts) / (ms * 1e-3), "permanents": sum(counts), "launches": 255, "ms": ms,
                          "note": "255 photon-subset inputs of |1^8,0^8>, 16 modes (SURVEY.md 8d C4), device-resident"}
    # BASELINE config 4 end to end (16 modes, 8 photons, imperfect source + detector loss): native source statistics,
    # distributions combined and sampled on the device, detector model vectorised on the host (one cold run)
    try:
        import lightworks_b200 as lb

        circ16 = lb.CompiledCircuitView(O.random_unitary(16, 4), 16)
        t0 = time.perf_counter()
        counts = lb.sample_source_outputs(circ16, [1] * 8 + [0] * 8, 1_000_000, lb.PermanentBackend(), purity=0.99,
                                          brightness=0.9, indistinguishability=0.95, detector_efficiency=0.9, seed=1)
        dt = time.perf_counter() - t0
        out["c4_noisy_sampler"] = {"samples": 1_000_000, "seconds": dt, "samples_per_s": 1e6 / dt,
                                   "distinct_outcomes": len(counts),
                                   "note": "Sampler(16-mode Haar seed 4, |1^8,0^8>, Source(purity 0.99, brightness 0.9, "
                                           "indistinguishability 0.95), Detector(efficiency 0.9), sampling_mode input): "
                                           "26 068 annotated inputs, 13.0 M output states, lightworks_b200."
                                           "sample_source_outputs; the reference needs 131 s for the source statistics alone"}
    except Exception as e:
        out["c4_noisy_sampler"] = {"error": repr(e)}
    U60 = O.random_unitary(60, 5)  # noqa: N806
    for n in (24, 28):
        with torch.cuda.stream(stream):
            A = torch.from_numpy(U60[:n, :n].copy()).cuda()  # noqa: N806
            o2 = torch.empty(2, dtype=torch.float64, device="cuda")
        ms = t_ms(lambda: ctx.perm_single_dev(A, n, o2))
        out[f"single_perm_n{n}"] = {"permanents_per_s": 1e3 / ms, "ms": ms, "fp64_frac": flops_per_perm(n) / (ms * 1e-3) / peak}
    return out


if __name__ == "__main__":
    sys.exit(main())
