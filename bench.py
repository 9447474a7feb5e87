#!/usr/bin/env python
"""bench.py — sub-system Newton solves/sec on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, host cores

Workload (config.workload): BASELINE.json configs[1] — 2^20 independent synthetic triangle
clusters per GPU, half K1 (distance+distance: ZeroFixedPoints / TwoFixedPointsDistance shapes),
half K5 (angle + unit normal), one solve2D-equivalent each (2 seeds, reference defaults) + root
selection (+ line reconstruction for K5).  A step = one pass over that batch = one K1 launch +
one K5 launch.  Weak scaling: rank r owns instances [r*2^19, (r+1)*2^19) of each kind's index
space; no data-path collective (NCCL only carries the barrier and the max-over-ranks of the time).

`value`   whole-job solves/s with the batch resident in HBM; per-launch CUDA events on the
          launching stream, summed over the steps, max over ranks; L2 flushed between steps.
`e2e`     the same metric through the host-buffer C-ABI calls (gcs_b200_solve_host_async per kind
          + gcs_b200_wait): pinned HOST inputs and outputs, H2D + kernels + D2H inside the region.
`bit_identical` (contracted variants, the default) the same step with the library-default kernels
          (bit-identical to the reference arithmetic), timed the same way in the same run, and the
          contract between the two checked on the whole batch (a failed check makes the line report
          the bit-identical kernels and name the violation: no number from unverified results).
`roofline` the dominant kernel (K1): algorithmic FP64 flops (work model of DESIGN.md section 3,
          from the MEASURED iteration counts) / its event-timed duration, against the DFMA peak
          measured live by gcs_b200_fp64_probe (MEASURED_PEAKS.json has no FP64 figure); the HBM
          side is reported beside it against MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`  the reference's own solve2D + primitives + heuristics sources
          (oracle/_ref, built from /root/reference against stand-in Eigen/autodiff headers) where
          that library is present, else the restated C oracle; all host threads, bounded sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1 << 20
METRIC = "subsystem Newton solves/sec"
UNIT = "solves/s"
WORKLOAD = ("configs[1]: 2^20 synthetic triangle clusters per GPU (2^19 K1 distance-distance + 2^19 K5 "
            "angle-normal), 2 seeds each, FP64")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=5,
                    help="5 (default) contracted arithmetic: discrete outputs identical to the reference, coordinates to 1e-9 "
                         "(6 static / 7 sorted); 0 the library default = bit-identical kernels (1 static, 2 refill, "
                         "3 sorted, 4 pair).  With a contracted variant the bit-identical default is timed beside it.")
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="solves per GPU (default 2^20)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "multistart8", "sweep64m"],
                    help="configs1 = the bench line (BASELINE configs[1]); multistart8 = configs[2] (2^20 K1 x 8 seeds per "
                         "GPU, weak); sweep64m = configs[4] (2^26 perturbed K1 instances generated on the device, sharded "
                         "over the ranks, strong)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the last committed `ncu --set full`
    capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler:
    """Polls NVML for SM clock + throttle reasons while a region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period = float(os.environ.get("GCS_BENCH_CLOCK_PERIOD", "0.001"))
        self._stop = threading.Event()
        self._in_region = False
        self.region = []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {}
        if nv:
            for k in dir(nv):
                if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                    v = getattr(nv, k)
                    if isinstance(v, int) and v not in (0,):
                        names.setdefault(v, k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            if nv:
                try:
                    mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append(mhz)
                    if self._in_region:
                        self.region.append(mhz)
                        for bit, nm in names.items():
                            if r & bit and nm not in ("GpuIdle", "None", "All"):
                                self.reasons.add(nm)
                except Exception:
                    pass
            time.sleep(self.period)

    def start(self):
        self.t.start()

    def enter(self):
        self._in_region = True

    def leave(self):
        self._in_region = False

    def stop(self):
        self._stop.set()
        self.t.join(timeout=1.0)
        src = self.region if self.region else self.samples
        return {
            "sm_mhz": float(np.median(src)) if src else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples_in_timed_region": len(self.region),
        }


def make_batches(synth, n, rank):
    half = n // 2
    first = rank * half
    return [synth.make_pp(half, first=first), synth.make_ang(half, first=first)]


# --------------------------------------------------------------------------------------------
# CPU leg: the reference's own sources where oracle/_ref exists, else the restated oracle
# --------------------------------------------------------------------------------------------
def cpu_backend():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    g.build_oracle()
    try:
        import ref_lib as R
        if R.available():
            R.load()
            return ("reference", lambda b, t: R.solve_batch(b, count_iters=False, threads=t), R.load().gcs_ref_max_threads(),
                    "the reference's solve2D / equation primitives / heuristics sources compiled from /root/reference "
                    "(oracle/_ref; Eigen and autodiff are the stand-in headers of oracle/ref_shim)")
    except Exception:
        pass
    import oracle_lib as O
    return ("port", lambda b, t: O.solve(b, t), O.max_threads(),
            "restated C oracle (oracle/gcs_oracle.c): oracle/_ref is not present on this box")


def cpu_rate(gcs, seconds, threads=0):
    """Solves/s of the CPU path on a bounded sample of the same workload (same generators, same
    K1/K5 mix), repeated until about `seconds` of wall time have been spent."""
    synth = gcs.synth
    kind, solve, cores, what = cpu_backend()
    if threads < 1:
        # every core this process may run on (torchrun exports OMP_NUM_THREADS=1; the CPU arm is
        # meant to use the whole host)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = threads
    m = 1 << 15
    bs = [synth.make_pp(m).alloc_outputs(), synth.make_ang(m).alloc_outputs()]
    for b in bs:  # warm the threads and the pages
        solve(b, threads)
    done, t_total, reps = 0, 0.0, 0
    while t_total < seconds or reps < 2:
        t0 = time.perf_counter()
        for b in bs:
            solve(b, threads)
        t_total += time.perf_counter() - t0
        done += 2 * m
        reps += 1
    sample = f"{reps} x ({m} K1 + {m} K5) solves of the bench generators, {cores} threads, {t_total:.1f} s; {what}"
    return done / t_total, cores, kind, sample


def run_reference(args, rank):
    if rank != 0:
        return
    gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
    n_steps = max(args.steps, 1)
    per_step = min(max(args.cpu_seconds / (n_steps + args.warmup), 0.25), 20.0)
    rates = []
    desc = cores = kind = None
    for i in range(args.warmup + n_steps):
        r, cores, kind, desc = cpu_rate(gcs, per_step)
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU path on the box's host cores, rank 0 only; each step is a bounded sample"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------
# the other BASELINE configs that are benchmarks (not the driver's bench line)
# --------------------------------------------------------------------------------------------
def run_extra(args, rank, local_rank, world):
    import ctypes as C
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
    capi, synth = gcs.capi, gcs.synth
    shard = importlib.import_module("2d_geometry_constraint_solver_b200.shard")
    capi.init([local_rank])
    lib = capi.load()
    warmup = max(args.warmup, 3)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if args.workload == "multistart8":
        n_total, scaling = (1 << 20) * world, "weak"
        n = 1 << 20
        hb = synth.make_pp(n, first=rank * n, n_seeds=8)
        hb.variant = args.variant
        db = capi.DeviceBatch(hb, dev, want_cand=False, variant=args.variant)
        name = "configs[2]: 2^20 K1 clusters per GPU x 8 initial guesses, orientation-based root selection"
        gen = None
    else:
        n_total, scaling = 1 << 26, "strong"
        lo, hi = shard.shard_range(n_total, rank, world)
        n = hi - lo
        hb = synth.make_pp(1)  # descriptor template; the columns are generated on the device
        db = capi.DeviceBatch.empty(capi.KIND_PP, 2, n, dev, variant=args.variant)
        ptrs = (C.c_void_p * 6)(*[c.data_ptr() for c in db.cols])

        def gen():
            capi.check(lib.gcs_b200_synth_pp(local_rank, C.c_void_p(stream.cuda_stream), synth.BASE_SEED, lo, n, 4096, ptrs,
                                             C.c_void_p(db.code.data_ptr())), "gcs_b200_synth_pp")
        gen()
        name = ("configs[4]: parametric sweep, 2^26 K1 instances = 4096 base clusters x 16384 perturbations (+-5% on ra, rb, d), "
                "generated on the device, sharded by index over the ranks")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        flush.fill_(1)
        db.solve()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = lib.gcs_b200_launch_count()
    barrier()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        evs[k][0].record(stream)
        db.solve()
        evs[k][1].record(stream)
    barrier()
    launches = lib.gcs_b200_launch_count() - l0
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    t = torch.tensor([float(ms.sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_total * args.steps / (total_ms * 1e-3)
    # work of this rank's launch from its measured iteration counts (device-side reduction)
    it = db.iters
    w = float((it.to(torch.int64) + 1).sum().item()) * synth.F_EVAL[1] + n * synth.F_SELECT[1]
    ach = w / (float(ms.mean()) * 1e-3) / 1e12
    conv = float(db.converged.to(torch.float32).mean().item())
    # contracted variants: the same batch through the bit-identical kernels, and the contract between
    # the two checked on this rank's whole batch (device-side comparison)
    contract = None
    if args.variant >= 5:
        if gen is None:
            db0 = capi.DeviceBatch(hb, dev, want_cand=False, variant=0)
        else:
            db0 = capi.DeviceBatch.empty(capi.KIND_PP, 2, n, dev, variant=0)
            for dst, src in zip(db0.cols, db.cols):
                dst.copy_(src)
            db0.code.copy_(db.code)
        db.solve()
        db0.solve()
        torch.cuda.synchronize(dev)
        same = bool(torch.equal(db.iters, db0.iters)) and bool(torch.equal(db.converged, db0.converged)) \
            and bool(torch.equal(db.root_index, db0.root_index))
        scale = torch.clamp(torch.stack([c.abs() for c in db.cols]).max(dim=0).values, min=1.0)
        worst = max(float(((x - y).abs() / torch.maximum(scale, y.abs())).max().item()) for x, y in zip(db.out, db0.out))
        ok = torch.tensor([1 if (same and worst <= 1e-9) else 0], device=dev, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            # never report a number from results that failed their check: rerun this workload on the
            # bit-identical kernels (every rank takes this branch together)
            print(f"[bench] CONTRACT VIOLATED by variant {args.variant} (discrete outputs equal = {same}, max relative error = "
                  f"{worst:.3e}); rerunning on the bit-identical kernels", file=sys.stderr, flush=True)
            if world > 1:
                dist.destroy_process_group()
            args.variant = 0
            return run_extra(args, rank, local_rank, world)
        contract = {"iters_flags_roots_equal": same, "max_rel_coordinate_error": worst, "tolerance": 1e-9, "solves_compared": int(n)}
        del db0, scale
    # end to end: generation (sweep) or pinned H2D (multi-start), solve, results back to pinned host memory
    e2e_steps = max(3, min(args.steps, 10))
    if gen is None:
        keep = []
        slab = torch.empty((6, n), dtype=torch.float64, pin_memory=True); slab.numpy()[...] = np.stack(hb.cols)
        code = torch.empty(n, dtype=torch.uint8, pin_memory=True); code.numpy()[...] = hb.code
        oslab = torch.empty((2, n), dtype=torch.float64, pin_memory=True)
        its = torch.empty((8, n), dtype=torch.int16, pin_memory=True)
        cvs = torch.empty((8, n), dtype=torch.uint8, pin_memory=True)
        root = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        keep += [slab, code, oslab, its, cvs, root]
        eb = capi.HostBatch(1, 8, [slab.numpy()[c] for c in range(6)], code.numpy(), None, args.variant, want_cand=False)
        eb.out = [oslab.numpy()[c] for c in range(2)]
        eb.iters, eb.converged, eb.root_index, eb.cand = its.numpy(), cvs.numpy(), root.numpy(), None

        def e2e_step():
            capi.solve_host(eb, local_rank)
        h2d, d2h = 6 * 8 * n + n, 2 * 8 * n + 8 * 3 * n + n
    else:
        outs = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        roots = torch.empty(n, dtype=torch.uint8, pin_memory=True)

        def e2e_step():
            gen()
            db.solve()
            for o, dcol in zip(outs, db.out):
                o.copy_(dcol, non_blocking=True)
            roots.copy_(db.root_index, non_blocking=True)
            torch.cuda.synchronize(dev)
        h2d, d2h = 0, 17 * n
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    et = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    if rank == 0:
        dfma = lib.gcs_b200_fp64_probe(local_rank, 0)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "solves_total": n_total, "solves_rank0": n, "l2": "256 MiB flush write between timed steps",
                       "converged_fraction_rank0": conv, "mean_iters_per_seed_rank0": float(it.to(torch.float32).mean().item()),
                       "variant": args.variant, "contract_check_rank0": contract},
            "e2e": {"value": n_total * e2e_steps / float(et.item()), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "gcs_b200_solve_host (pinned)" if gen is None else "gcs_b200_synth_pp + gcs_b200_solve + D2H of (x, y, root) to pinned memory"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": lib.gcs_b200_kernel_name(1, 8 if gen is None else 2, args.variant).decode(), "bound": "fp64",
                         "achieved": ach, "peak": dfma, "unit": "TFLOP/s", "frac": ach / dfma if dfma > 0 else None, "traffic": None,
                         "algorithmic_flops_per_launch_rank0": w},
            "cpu_baseline": None,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
def _quiet_stdout():
    """Library chatter on file descriptor 1 (NCCL prints its version there on some boxes) goes to
    stderr; the JSON line is the only thing written to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    args = parse()
    _quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload != "configs1":
        run_extra(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
    capi, synth = gcs.capi, gcs.synth
    capi.init([local_rank])
    lib = capi.load()

    n = args.n
    warmup = max(args.warmup, 3)
    host = make_batches(synth, n, rank)
    for h in host:
        h.variant = args.variant
    devb = [capi.DeviceBatch(h, dev, want_cand=False, variant=args.variant) for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stream = torch.cuda.current_stream(dev)

    def step(events=None):
        if events:
            events[0].record(stream)
        devb[0].solve()
        if events:
            events[1].record(stream)
        devb[1].solve()
        if events:
            events[2].record(stream)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        flush.fill_(1)
        step()
    barrier()

    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    l0 = lib.gcs_b200_launch_count()
    sampler.enter()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # evict the batch from L2 between timed steps (outside the events)
        step(evs[k])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.gcs_b200_launch_count() - l0

    ms_k1 = np.array([e[0].elapsed_time(e[1]) for e in evs])
    ms_k5 = np.array([e[1].elapsed_time(e[2]) for e in evs])
    ms_step = ms_k1 + ms_k5
    t_sum = torch.tensor([float(ms_step.sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_sum, op=dist.ReduceOp.MAX)
    total_ms = float(t_sum.item())
    value = (n * world * args.steps) / (total_ms * 1e-3)

    # ---- the same step with the two launches on two streams (the kinds are independent batches):
    #      one kernel's tail - the few warps that re-run a seed with the literal code, ~8 us of
    #      dependent FP64 latency each - is covered by the other kernel.  Reported beside `value`,
    #      which stays the sum of the per-launch times. ----
    side = torch.cuda.Stream(device=dev)
    fork = [torch.cuda.Event() for _ in range(args.steps)]
    join = [torch.cuda.Event() for _ in range(args.steps)]
    ev2 = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]

    def step2(k):
        ev2[k][0].record(stream)
        fork[k].record(stream)
        side.wait_event(fork[k])
        devb[0].solve(stream)
        devb[1].solve(side)
        join[k].record(side)
        stream.wait_event(join[k])
        ev2[k][1].record(stream)

    for k in range(min(3, args.steps)):
        flush.fill_(1)
        step2(k)
    barrier()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        step2(k)
    barrier()
    ms2 = np.array([e[0].elapsed_time(e[1]) for e in ev2])
    t2_sum = torch.tensor([float(ms2.sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2_sum, op=dist.ReduceOp.MAX)
    two_stream = {"value": (n * world * args.steps) / (float(t2_sum.item()) * 1e-3), "unit": UNIT,
                  "ms_per_step": float(t2_sum.item()) / args.steps,
                  "note": "K1 and K5 launches of a step enqueued on two streams, CUDA events from fork to join on the launching stream"}

    # ---- the bit-identical kernels on the same batches, timed the same way, and the contract
    #      between the two checked on the full batch (contracted variants only) ----
    bit_identical = None
    violation = None
    if args.variant >= 5:
        devb0 = [capi.DeviceBatch(h, dev, want_cand=False, variant=0) for h in host]
        for _ in range(warmup):
            flush.fill_(1)
            for d in devb0:
                d.solve()
        barrier()
        evs0 = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            evs0[k][0].record(stream)
            devb0[0].solve()
            evs0[k][1].record(stream)
            devb0[1].solve()
            evs0[k][2].record(stream)
        barrier()
        ms0_k1 = np.array([e[0].elapsed_time(e[1]) for e in evs0])
        ms0_k5 = np.array([e[1].elapsed_time(e[2]) for e in evs0])
        t0_sum = torch.tensor([float((ms0_k1 + ms0_k5).sum())], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t0_sum, op=dist.ReduceOp.MAX)
        same, worst = True, 0.0
        for a, b, h in zip(devb, devb0, host):
            same = same and bool(torch.equal(a.iters, b.iters)) and bool(torch.equal(a.converged, b.converged)) \
                and bool(torch.equal(a.root_index, b.root_index))
            scale = torch.clamp(torch.stack([c.abs() for c in a.cols]).max(dim=0).values, min=1.0)
            for x, y in zip(a.out, b.out):
                worst = max(worst, float(((x - y).abs() / torch.maximum(scale, y.abs())).max().item()))
        if os.environ.get("GCS_BENCH_TEST_VIOLATION"):  # exercises the fallback below
            same = False
        ok = torch.tensor([1 if (same and worst <= 1e-9) else 0], device=dev, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            # Never report a number from results that failed their check: the line falls back to the
            # bit-identical kernels (timed above, on the same batches) and says so.
            print(f"[bench] CONTRACT VIOLATED by variant {args.variant} on rank {rank}'s view (discrete outputs equal = {same}, "
                  f"max relative error = {worst:.3e}); reporting the bit-identical kernels instead", file=sys.stderr, flush=True)
            violation = {"variant": args.variant, "iters_flags_roots_equal_rank": same, "max_rel_coordinate_error_rank": worst}
            args.variant = 0
            two_stream = None
            devb, ms_k1, ms_k5 = devb0, ms0_k1, ms0_k5
            total_ms = float(t0_sum.item())
            value = (n * world * args.steps) / (total_ms * 1e-3)
            for h in host:
                h.variant = 0
        bit_identical = {
            "variant": "default (newton_sorted_kernel, literal Householder QR, no contraction)",
            "value": (n * world * args.steps) / (float(t0_sum.item()) * 1e-3), "unit": UNIT,
            "ms_per_step": float(t0_sum.item()) / args.steps,
            "k1_launch_ms": float(ms0_k1.mean()), "k5_launch_ms": float(ms0_k5.mean()),
            "contract_check_rank0": {"iters_flags_roots_equal": same, "max_rel_coordinate_error": worst, "tolerance": 1e-9,
                                     "solves_compared": int(sum(d.n for d in devb))},
        }
        if violation is None:
            del devb0

    # ---- iteration histogram -> algorithmic work of the dominant kernel ----
    it1 = devb[0].iters.cpu().numpy()
    it5 = devb[1].iters.cpu().numpy()
    w_k1 = synth.algorithmic_flops(1, it1)
    w_k5 = synth.algorithmic_flops(5, it5)
    k1_ms = float(np.mean(ms_k1))
    k5_ms = float(np.mean(ms_k5))
    b_k1 = devb[0].algorithmic_bytes()
    b_k5 = devb[1].algorithmic_bytes()

    # ---- end to end through the host-buffer C-ABI calls, pinned host memory ----
    def pin_like(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
        v = t.numpy()
        v[...] = a
        return t, v

    keep = []
    e2e_batches = []
    for h in host:
        # one pinned [columns][n] slab per direction: the library moves such slabs with strided copies
        t, slab = pin_like(np.stack(h.cols)); keep.append(t)
        cols = [slab[c] for c in range(slab.shape[0])]
        t, code = pin_like(h.code); keep.append(t)
        hb = capi.HostBatch(h.kind, h.n_seeds, cols, code, None, args.variant, want_cand=False)
        m = hb.n
        t, oslab = pin_like(np.zeros((capi.OUT_COLS[h.kind], m))); keep.append(t)
        hb.out = [oslab[c] for c in range(oslab.shape[0])]
        t, hb.iters = pin_like(np.zeros((h.n_seeds, m), np.int16)); keep.append(t)
        t, hb.converged = pin_like(np.zeros((h.n_seeds, m), np.uint8)); keep.append(t)
        t, hb.root_index = pin_like(np.zeros(m, np.uint8)); keep.append(t)
        hb.cand = None
        e2e_batches.append(hb)
    h2d = sum(capi.IN_COLS[b.kind] * 8 * b.n + b.n for b in e2e_batches)
    d2h = sum(capi.OUT_COLS[b.kind] * 8 * b.n + b.n_seeds * 3 * b.n + b.n for b in e2e_batches)
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        for b in e2e_batches:
            capi.solve_host_async(b, local_rank)
        capi.wait(local_rank)

    for _ in range(3):
        e2e_step()
    l1 = lib.gcs_b200_launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.leave()
    e2e_launches = lib.gcs_b200_launch_count() - l1
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n * world * e2e_steps / float(e2e_t.item())
    # the e2e results must equal the device-resident ones (same inputs)
    assert np.array_equal(e2e_batches[0].iters, it1) and np.array_equal(e2e_batches[1].iters, it5)
    assert np.array_equal(e2e_batches[0].out[0], devb[0].out[0].cpu().numpy())

    clocks = sampler.stop()

    # ---- peaks (after the timed regions: the probe heats the chip) ----
    dfma_peak = lib.gcs_b200_fp64_probe(local_rank, 0)
    mix_peak = lib.gcs_b200_fp64_probe(local_rank, 1)
    peaks, peak_src = load_peaks()
    ach_tf = w_k1 / (k1_ms * 1e-3) / 1e12
    traffic = load_traffic()
    vname = {0: "default", 1: "static", 2: "refill", 3: "sorted", 4: "pair", 5: "contracted", 6: "contracted-static",
             7: "contracted-sorted"}[args.variant]
    rerun_stats = None
    if args.variant >= 5:
        import ctypes as C
        st2 = (C.c_uint64 * 2)()
        lib.gcs_b200_contracted_stats(local_rank, st2, 1)
        step()
        lib.gcs_b200_contracted_stats(local_rank, st2, 1)
        rerun_stats = {"runs_per_step": int(sum(d.n * d.n_seeds for d in devb)), "redone_by_run_guards": int(st2[0]),
                       "redone_by_selection_guard": int(st2[1])}
    kernel_name = lib.gcs_b200_kernel_name(1, 2, args.variant).decode() if hasattr(lib, "gcs_b200_kernel_name") else vname
    roofline = {
        "kernel": kernel_name,
        "bound": "fp64",
        "bound_note": "FP64 CUDA-core pipe (no tensor-core work on this path); the HBM side is under 'hbm'",
        "achieved": ach_tf, "peak": dfma_peak, "unit": "TFLOP/s", "frac": ach_tf / dfma_peak if dfma_peak > 0 else None,
        "peak_source": "measured live: gcs_b200_fp64_probe DFMA micro-benchmark (FMA = 2 flops); MEASURED_PEAKS.json has no FP64 entry",
        "non_fma_peak": mix_peak,
        "frac_of_non_fma_peak": ach_tf / mix_peak if mix_peak > 0 else None,
        "traffic": traffic.get(kernel_name, {}).get("dram_bytes_per_launch"),
        "traffic_source": traffic.get(kernel_name, {}).get("source"),
        "algorithmic_flops_per_launch": w_k1,
        "algorithmic_flops_note": ("SURVEY.md 8d work model: (iters+1) * 64 flops per seed + selection, from the measured iteration "
                                   "counts - the reference algorithm's work (Householder QR counted at 44 flops per update).  The "
                                   "contracted kernels reach the same iterates with a closed-form solve (~33 executed flops per update, "
                                   "FMA = 2), so for them `achieved` is a rate of reference-algorithm work, not of executed flops; "
                                   "executed flops per launch are in profiles/traffic.json" if args.variant >= 5 else
                                   "SURVEY.md 8d work model from the measured iteration counts"),
        "executed_flops_per_launch": traffic.get(kernel_name, {}).get("executed_flops_per_launch"),
        "flops_per_solve": w_k1 / devb[0].n,
        "mean_iters_per_seed": float(it1.mean()),
        "launch_ms": k1_ms,
        "share_of_step": k1_ms / (k1_ms + k5_ms),
        "hbm": {"achieved": b_k1 / (k1_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": b_k1 / (k1_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 1.0), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_k1},
        "second_kernel": {"kernel": "K5", "launch_ms": k5_ms, "achieved_tflops": w_k5 / (k5_ms * 1e-3) / 1e12,
                          "hbm_gbs": b_k5 / (k5_ms * 1e-3) / 1e9, "mean_iters_per_seed": float(it5.mean()),
                          "algorithmic_flops_per_launch": w_k5, "algorithmic_bytes_per_launch": b_k5},
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, kind, desc = cpu_rate(gcs, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "solves_per_gpu": n, "l2": "256 MiB flush write between timed steps",
                       "variant": vname,
                       "parity_class": ("contract: iteration counts, convergence flags and root indices identical to the reference, "
                                        "coordinates within 1e-9 relative (checked in this run against the bit-identical kernels)"
                                        if args.variant >= 5 else "bit-identical to the reference arithmetic"),
                       "literal_reruns": rerun_stats,
                       "contract_violation": violation,
                       "timing": "per-launch CUDA events on the launching stream, sum over steps, max over ranks",
                       "wall_s_timed_region_incl_flush": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "gpu_launches": int(e2e_launches),
                    "api": "gcs_b200_solve_host_async x2 + gcs_b200_wait (pinned host buffers, wall clock incl. copies)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "two_stream_step": two_stream,
            "bit_identical": bit_identical,
            "cpu_baseline": cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
