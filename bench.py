#!/usr/bin/env python3
"""bench.py -- headline measurement of the hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-coeffs 24]

A "step" = one fold-high (sumcheck fold / extrapolate_line) over one multilinear
of 2^24 GF(2^128) input coefficients per GPU (256 MiB read, 128 MiB written); the steps rotate over 4
resident multilinears (1 GiB, far larger than the 126 MB L2, so no flush is needed).  `value` =
coefficients/s over all ranks with inputs resident in HBM; `e2e` = the same metric through the plugin
call with HOST buffers (copy_h2d + fold + copy_d2h inside the timed region).  Next to the headline:
  sumcheck_e2e     whole sumchecks through the trait calls with ONE upload (what the ComputeLayer contract
                   exercises): the fold chain of one 2^24 multilinear and a bivariate-product sumcheck
                   (m = 8, n = 20, round evaluations through traced accumulate_kernels), each with a CPU arm
  ntt              BASELINE config #2 (S1/S2/S3) with the CPU arm of S1
  keccak_replay    the compiled op-sequence replay of the keccak example at 2^18 permutations incl. the
                   witness upload and the univariate-skip round, with a CPU arm assembled from bounded samples
  sharded_sumcheck (N > 1) the low-variable-sharded bivariate sumcheck over NCCL, bit-exact vs the oracle
N > 1: one process per GPU (torchrun), independent multilinears per rank (weak scaling, no data-path
collective in the headline); the only collective of the headline is the timing max-reduce.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sumcheck_fold_high_gf2_128_coeffs_per_s"
UNIT = "coeffs/s"
N_BUFFERS = 4


def bench_config(log_coeffs):
    """identical in both arms (`--impl ours` and `--impl reference`)"""
    return {"workload": f"fold-high (extrapolate_line) over BinaryField128b, 2^{log_coeffs} coefficients per GPU",
            "l2_policy": f"{N_BUFFERS} rotating multilinears of 2^{log_coeffs} coefficients (larger than L2); no flush"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def profiled_traffic(kernel="k_lerp_tma"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` summary of this same command (profiles/); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", f"r2_fold_{kernel}_ncu_full.csv")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", f"r1_fold_{kernel}_ncu_full.csv")
    try:
        rd = wr = None
        for line in open(path):
            p = line.strip().split(",")
            if len(p) >= 3 and p[0] == "dram__bytes_read.sum" and rd is None:
                rd = float(p[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[p[1]]
            if len(p) >= 3 and p[0] == "dram__bytes_write.sum" and wr is None:
                wr = float(p[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[p[1]]
        return None if rd is None or wr is None else {"bytes": rd + wr, "source": os.path.relpath(path, ROOT)}
    except Exception:
        return None


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank (and therefore the pinned host buffers it first-touches) to the CPUs next to its GPU,
    so that at N > 1 the host<->device copies of the e2e path do not cross the socket interconnect."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1 and 64 * w + b < n_cpu}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


class NvmlClocks:
    """In-process NVML reads (~0.1 ms each): the timed region of the fold bench lasts ~1.5 ms, shorter than
    nvidia-smi's sampling period, so the clocks and throttle reasons are read by the host thread WHILE the
    queued steps execute (after submission, before the synchronising event read)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, gpu_index):
        self.h = None
        self.sm, self.reasons = [], set()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def sample(self, n=1):
        if self.h is None:
            return
        for _ in range(n):
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                return

    def result(self):
        if self.h is None or not self.sm:
            return None
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(self.sm),
                "window": "NVML reads by the host thread while the timed steps execute"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[1]))
                mx = float(p[2])
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": "nvidia-smi -lms 100 around the timed region"}


class Ev:
    """CUDA events on the library's own stream (torch.cuda.Event would only see torch's stream)."""

    def __init__(self, hal):
        self.hal = hal
        self.a, self.b = C.c_void_p(), C.c_void_p()
        hal._check(hal._lib.b200_event_create(hal._ctx, C.byref(self.a)))
        hal._check(hal._lib.b200_event_create(hal._ctx, C.byref(self.b)))

    def start(self):
        self.hal._check(self.hal._lib.b200_event_record(self.hal._ctx, self.a))

    def record_stop(self):
        self.hal._check(self.hal._lib.b200_event_record(self.hal._ctx, self.b))

    def elapsed_ms(self):
        ms = C.c_float()
        self.hal._check(self.hal._lib.b200_event_elapsed_ms(self.hal._ctx, self.a, self.b, C.byref(ms)))
        return float(ms.value)

    def stop_ms(self):
        self.hal._check(self.hal._lib.b200_event_record(self.hal._ctx, self.b))
        ms = C.c_float()
        self.hal._check(self.hal._lib.b200_event_elapsed_ms(self.hal._ctx, self.a, self.b, C.byref(ms)))
        return float(ms.value)


def cpu_fold_baseline(log_coeffs, budget_s=12.0):
    """CPU arm: C restatement of the reference's fold loop (fold_left_lerp_inplace / extrapolate_line)
    with its AVX-512+GFNI multiply, all host threads, timed on a bounded sample (oracle/cpu_baseline.c)."""
    from oracle import binding as orc

    orc.lib()
    return orc.cpu_fold_parallel(log_coeffs, budget_s)


def cpu_univariate_baseline(rows_log2=21, columns=153, compositions=75, skip=7):
    """CPU arm of the univariate-skip round (oracle/cpu_univariate.c: table-driven port, all host threads) on a
    bounded sample of the bench shape: the same columns/constraints on 2^rows_log2 rows."""
    from oracle import binding as orc

    words = 1 << (rows_log2 - 7)
    cols = [orc.rand_b128(j, words) for j in range(columns)]
    m = columns
    comps = []
    for c in range(compositions):  # the same chi constraints as the GPU arm: out + b0 + b1 * b2 + b2
        batch, xy = c // 25, c % 25
        x, y = xy % 5, xy // 5
        b0, b1, b2 = (75 + 25 * batch + (x + k) % 5 + 5 * y for k in range(3))
        comps.append([("var", c % m), ("var", b0 % m), ("add", 0, 1), ("var", b1 % m), ("var", b2 % m), ("mul", 3, 4), ("add", 2, 5), ("add", 6, 4)])
    eq = orc.rand_b128(999, 1 << (rows_log2 - skip))
    cores = len(os.sched_getaffinity(0))
    best = min(orc.cpu_univariate_b1(cols, rows_log2, skip, eq, comps, 1 << skip, cores)[1] for _ in range(2))
    return {"value": columns * (1 << (rows_log2 - skip)) / best, "unit": "sub-cube columns/s", "cores": cores, "kind": "port",
            "sample": f"{columns} B1 columns x 2^{rows_log2} rows, {compositions} degree-2 constraints, skip {skip}: {best * 1e3:.1f} ms "
                      "(table-driven C restatement, pthreads; includes the table setup)"}


def cpu_keccak_replay_baseline():
    """CPU arm of the keccak replay, assembled from bounded samples of the same phases (each a threaded C restatement
    of the reference's GFNI path, oracle/cpu_baseline.c + cpu_univariate.c) and scaled to the 2^18-permutation sizes;
    the FRI folds and ring-switch eq-indicators (small on both arms) are left out."""
    from oracle import binding as orc

    cores = len(os.sched_getaffinity(0))
    uni = cpu_univariate_baseline(rows_log2=21)               # 153 columns x 2^21 rows, skip 7  -> x 64
    uni_ms = 153 * (1 << (27 - 7)) / uni["value"] * 1e3
    ntt = orc.cpu_ntt_parallel(6, 20, 1, 3.0, d=24)           # 2^26 coefficients, 19 layers     -> x 16 x 23/19
    ntt_ms = ntt["ms_per_transform"] * 16 * 23 / 19
    zc = orc.cpu_chi_zerocheck_parallel(75, 78, 16, 3.0)      # 153 multilinears of 2^16         -> x 16
    zc_ms = zc["ms_per_sumcheck"] * 16
    pi = orc.cpu_bivariate_sumcheck_parallel(40, 18, 20, 3.0)  # 40 multilinears of 2^18, 20 pairs -> x 5 x 4
    pi_ms = pi["ms_per_sumcheck"] * 20
    return {"total_ms": uni_ms + ntt_ms + zc_ms + pi_ms, "cores": cores, "kind": "port",
            "phases": {"zerocheck_univariate_skip_round": {"ms": uni_ms}, "commit_rs_encode_ntt": {"ms": ntt_ms},
                       "zerocheck_rounds": {"ms": zc_ms}, "piop_bivariate_sumcheck": {"ms": pi_ms}},
            "sample": "extrapolated from bounded samples: univariate round at 2^21 rows (x64), NTT at 2^26 coefficients / 19 layers "
                      "(x16 x 23/19), chi zerocheck at 2^16 (x16), bivariate sumcheck at 40 x 2^18 / 20 pairs (x20); no witness upload, "
                      "no FRI / ring-switch phases; C restatements of the reference GFNI path, not the Rust binary"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    ms = []
    base = None
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        # a step = a bounded sample of the workload, sized so that the whole run ends within a few minutes at any K
        base = cpu_fold_baseline(args.log_coeffs, budget_s=min(1.5, 120.0 / max(args.warmup + args.steps, 1)))
        if s >= args.warmup:
            ms.append((time.perf_counter() - t0) * 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u128 (GF(2^128) tower)", "data": "synthetic",
            "config": bench_config(args.log_coeffs),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--log-coeffs", type=int, default=24)
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-keccak", action="store_true")
    ap.add_argument("--no-compiled-cfg3", action="store_true", help="skip the compiled-host runs of config #3 (under ncu the persistent "
                    "kernel cannot get its challenges: the profiler serialises the launch call)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import binius_b200
    from binius_b200 import NTTShape

    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # rank 0 prints ONE line on stdout: no NCCL version banner before it
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hal = binius_b200.B200Layer(local_rank)
    peak, peak_src = peaks()

    # ---- synthetic multilinear: 2^log_coeffs B128 coefficients (SplitMix64, seed = rank) ----------
    n_in = 1 << args.log_coeffs
    half = n_in // 2
    rng = np.random.default_rng(rank)
    host = rng.integers(0, 1 << 63, size=(n_in, 2), dtype=np.int64).astype(np.uint64)
    ph = C.c_void_p()
    hal._check(hal._lib.b200_host_alloc(hal._ctx, n_in * 16, C.byref(ph)))
    pinned = np.ctypeslib.as_array((C.c_uint64 * (2 * n_in)).from_address(ph.value)).reshape(n_in, 2)
    pinned[:] = host
    arena = hal.dev_alloc(N_BUFFERS * n_in)  # 4 resident multilinears (1 GiB at 2^24): the steps rotate over them
    bufs = [arena.slice(k * n_in, (k + 1) * n_in) for k in range(N_BUFFERS)]
    for b in bufs:
        hal._check(hal._lib.b200_copy_h2d(hal._ctx, ph.value, b.ptr, n_in))
    dev = bufs[0]
    z = 0x2E895399AF449ACE499596F6E5FCCAFA
    zs = (C.c_uint64 * 2)(z & (2**64 - 1), z >> 64)
    step_no = [0]

    def fold_step():
        b = bufs[step_no[0] % N_BUFFERS]
        step_no[0] += 1
        hal._check(hal._lib.b200_extrapolate_line(hal._ctx, b.ptr, half, b.ptr + 16 * half, half, zs))
        hal._check(hal._lib.b200_flush(hal._ctx))  # one launch per step (consecutive folds would otherwise batch)

    def barrier():
        hal.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    ev = Ev(hal)
    for _ in range(args.warmup):
        fold_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    l0 = hal.launch_count()
    t0 = time.time()
    nvml = NvmlClocks(local_rank)
    ev.start()
    for _ in range(args.steps):
        fold_step()
    ev.record_stop()
    nvml.sample(6)  # the queued steps are executing now
    ms_total = ev.elapsed_ms()
    barrier()
    t1 = time.time()
    launches = hal.launch_count() - l0

    # per-launch duration of the dominant kernel (k_lerp_tma), CUDA events around single launches
    kern_ms = []
    for _ in range(min(args.steps, 10)):
        hal._check(hal._lib.b200_sync(hal._ctx))
        ev.start()
        fold_step()
        kern_ms.append(ev.stop_ms())
    kern_ms_avg = float(np.mean(kern_ms))

    # ---- e2e: HOST buffers through the plugin call (b200_extrapolate_line_host: pipelined H2D of both
    #      halves, fold kernel, D2H of the folded half), pinned host memory, wall clock around the call
    e2e_steps = max(3, min(args.steps, 5))
    zint = z
    h_lo, h_hi = pinned[:half], pinned[half:]
    hal.extrapolate_line_host(h_lo, h_hi, zint)  # warm-up (stream/event creation, scratch allocation)
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        hal.extrapolate_line_host(h_lo, h_hi, zint)
    barrier()
    e2e_ms = (time.perf_counter() - te0) * 1e3 / e2e_steps
    clocks = sampler.stop(t0, t1)
    if nvml.result() is not None:
        clocks = dict(nvml.result(), nvidia_smi=clocks)

    # ---- sumcheck chain: what the prover does with one multilinear over a whole sumcheck -- upload once,
    #      fold-high log_coeffs times (2^24, 2^23, ..., 2 coefficients, a fresh challenge per round), read the
    #      final evaluation back.  Same ops through the same plugin calls; the PCIe copy is paid once.
    def chain_device():
        n = n_in
        r = 0
        while n >= 2:
            h = n // 2
            zr = (C.c_uint64 * 2)((z + r) & (2**64 - 1), z >> 64)
            hal._check(hal._lib.b200_extrapolate_line(hal._ctx, dev.ptr, h, dev.ptr + 16 * h, h, zr))
            n, r = h, r + 1

    chain_coeffs = 2 * n_in - 2
    for _ in range(2):
        chain_device()
    hal.sync()
    lc0 = hal.launch_count()
    ev.start()
    chain_reps = 5
    for _ in range(chain_reps):
        chain_device()
    chain_ms = ev.stop_ms() / chain_reps
    chain_launches = (hal.launch_count() - lc0) // chain_reps
    out1 = np.zeros((1, 2), np.uint64)
    barrier()
    tc0 = time.perf_counter()
    for _ in range(3):
        hal._check(hal._lib.b200_copy_h2d(hal._ctx, ph.value, dev.ptr, n_in))
        chain_device()
        hal._check(hal._lib.b200_copy_d2h(hal._ctx, dev.ptr, out1.ctypes.data, 1))
    barrier()
    chain_e2e_ms = (time.perf_counter() - tc0) * 1e3 / 3
    chain = {"coeffs_per_chain": chain_coeffs, "rounds": args.log_coeffs, "launches": int(chain_launches), "device_ms": chain_ms,
             "device_coeffs_per_s": chain_coeffs / (chain_ms * 1e-3), "e2e_ms": chain_e2e_ms,
             "e2e_coeffs_per_s": chain_coeffs / (chain_e2e_ms * 1e-3), "h2d_bytes": n_in * 16, "d2h_bytes": 16}

    # ---- whole bivariate-product sumcheck through the trait calls (v3::BivariateSumcheckProver's data plane,
    #      bivariate_product.rs:168-232, 303-408): upload m = 8 multilinears of 2^20 once, then 20 rounds of
    #      {calculate_round_evals = the traced accumulate_kernels closure, fold of every multilinear}, read the finals
    biv = None
    if not args.no_ntt:
        import random as _rb

        from binius_b200 import calculate_round_evals

        nvb, mb = 20, 8
        rb = _rb.Random(11)
        pairs_b = [((5 * c + 1) % mb, (3 * c + 2) % mb) for c in range(8)]
        h_mls = pinned[: mb << nvb]
        d_mls = bufs[1].slice(0, mb << nvb)
        finals = np.zeros((mb, 2), np.uint64)

        def biv_sumcheck(upload):
            if upload:
                hal._check(hal._lib.b200_copy_h2d(hal._ctx, h_mls.ctypes.data, d_mls.ptr, mb << nvb))
            cur = [d_mls.slice(t << nvb, (t + 1) << nvb) for t in range(mb)]
            for r in range(nvb):
                calculate_round_evals(hal, nvb - r, rb.getrandbits(128), cur, pairs_b)  # synchronous: the transcript needs the values
                ch = rb.getrandbits(128)

                def fold(ex):
                    for t in range(mb):
                        lo_, hi_ = cur[t].split_half_mut()
                        ex.extrapolate_line(lo_, hi_, ch)
                        cur[t] = lo_
                    return []

                hal.execute(fold)
            if upload:
                for t in range(mb):
                    hal._check(hal._lib.b200_copy_d2h(hal._ctx, cur[t].ptr, finals[t:].ctypes.data, 1))

        biv_sumcheck(True)
        barrier()
        lb0 = hal.launch_count()
        ev.start()
        for _ in range(3):
            biv_sumcheck(False)
        biv_dev_ms = ev.stop_ms() / 3
        biv_launches = (hal.launch_count() - lb0) // 3
        barrier()
        tb0 = time.perf_counter()
        for _ in range(3):
            biv_sumcheck(True)
        barrier()
        biv_e2e_ms = (time.perf_counter() - tb0) * 1e3 / 3
        biv = {"multilinears": mb, "n_vars": nvb, "compositions": len(pairs_b), "route": "accumulate_kernels traced in the library -> k_pair_tc",
               "launches": int(biv_launches), "device_ms": biv_dev_ms, "device_rounds_per_s": nvb / (biv_dev_ms * 1e-3), "e2e_ms": biv_e2e_ms,
               "e2e_rounds_per_s": nvb / (biv_e2e_ms * 1e-3), "h2d_bytes": (mb << nvb) * 16, "d2h_bytes": 16 * mb + 32 * nvb,
               "note": "device_ms includes the per-round synchronising fetch of the round values (Python mirror in the loop)"}

    # ---- NTT (BASELINE config #2): B32, 2^24 coefficients, S1/S2/S3 shapes ------------------------
    ntt_res = None
    if not args.no_ntt:
        ntt = binius_b200.B200AdditiveNTT(hal, 5, 24)
        n32 = 1 << 24
        ntt_res = {}
        for name, (lx, ly, lz, skip) in {"S1_rs_encode": (6, 18, 0, 1), "S2_single": (0, 24, 0, 0), "S3_batch": (0, 16, 8, 0)}.items():
            S = NTTShape(lx, ly, lz)
            for _ in range(2):
                ntt.forward_device(dev.ptr, 5, n32, S, 0, 0, skip)
            hal.sync()
            c0 = hal.launch_count()
            ev.start()
            reps = 5
            for _ in range(reps):
                ntt.forward_device(dev.ptr, 5, n32, S, 0, 0, skip)
            fwd_ms = ev.stop_ms() / reps
            passes = (hal.launch_count() - c0) // reps
            ev.start()
            for _ in range(reps):
                ntt.inverse_device(dev.ptr, 5, n32, S, 0, 0, skip)
            inv_ms = ev.stop_ms() / reps
            n_prod = (ly - skip) * n32 // 2
            ntt_res[name] = {"fwd_ms": fwd_ms, "inv_ms": inv_ms, "coeffs_per_s": n32 / (fwd_ms * 1e-3), "passes": passes,
                             "roofline_frac": (8 * n32 / (fwd_ms * 1e-3)) / (peak * 1e9),
                             "roofline": {"hbm_frac": (8 * n32 / (fwd_ms * 1e-3)) / (peak * 1e9), "b32_products_per_s": n_prod / (fwd_ms * 1e-3),
                                          "sm_cycles_per_product": fwd_ms * 1e-3 * 1.965e9 * 148 / n_prod,
                                          "limiting_pipes": "ALU + LSU (nibble/byte look-up tables: 8-19 ALU-pipe instructions and 4-8 shared-memory "
                                                            "gathers per product; profiles/r2_ntt_k_ntt_lut_ncu_full.csv)"}}

    # ---- sumcheck round (BASELINE config #3 shape, v3 bivariate variant): m = 8 multilinears, n = 20,
    #      8 compositions: round evaluations + fold of every multilinear; and one tensor expansion ------
    extras = None
    if not args.no_ntt:
        import random as _r

        rr = _r.Random(7)
        nv, m = 20, 8
        sub = [dev.slice(t << nv, (t + 1) << nv) for t in range(m)]  # 8 x 2^20 elements of a resident buffer
        pairs = [(rr.randrange(m), rr.randrange(m)) for _ in range(m)]
        alpha = rr.getrandbits(128)
        from binius_b200 import calculate_round_evals as _cre

        for _ in range(2):
            _cre(hal, nv, alpha, sub, pairs)
        ev.start()
        reps = 5
        for _ in range(reps):
            _cre(hal, nv, alpha, sub, pairs)  # the TRAIT route: accumulate_kernels closure traced by the library
        re_ms = ev.stop_ms() / reps
        ev.start()
        for _ in range(reps):
            hal.execute(lambda ex: list(ex.bivariate_round_evals(sub, nv, pairs, alpha)))
        re_fused_ms = ev.stop_ms() / reps
        k = 22
        coords = [rr.getrandbits(128) for _ in range(k)]
        te = dev.slice(0, 1 << k)
        hal.fill(te.slice(0, 1), 1)
        for _ in range(2):
            hal.execute(lambda ex: (ex.tensor_expand(0, coords, te), [])[1])
        ev.start()
        for _ in range(reps):
            hal.execute(lambda ex: (ex.tensor_expand(0, coords, te), [])[1])
        te_ms = ev.stop_ms() / reps
        extras = {"bivariate_round_evals_m8_n20": {"ms": re_ms, "route": "accumulate_kernels (traced)", "fused_entry_point_ms": re_fused_ms,
                                                   "products_per_s": 2 * len(pairs) * (1 << (nv - 1)) / (re_ms * 1e-3),
                                                   "hbm_frac": (16 * m * (1 << nv) / (re_ms * 1e-3)) / (peak * 1e9)},
                  "tensor_expand_k22": {"ms": te_ms, "elems_per_s": (1 << k) / (te_ms * 1e-3),
                                        "hbm_frac": (16 * (1 << k) / (te_ms * 1e-3)) / (peak * 1e9)}}

    # ---- BASELINE config #3: zerocheck round reduction of the u32_add circuit at 2^20 rows --
    #      5 multilinears of 18 variables after the univariate skip, eq-ind of 2^17, compositions
    #      (x+c)(y+c)+c-o and x+y+c-z (m3/src/gadgets/add.rs:71-76), HighToLow, 18 rounds of {round
    #      evaluations at 1 and infinity, fold of every multilinear, halving of the eq-indicator}
    cfg3 = None
    if not args.no_ntt:
        from binius_b200 import ArithCircuit as A
        from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear

        be = B200Backend(hal)
        nv3 = 18
        xv, yv, cin, cout, zv = (A.var(i) for i in range(5))
        comps3 = [(xv + cin) * (yv + cin) + cin - cout, xv + yv + cin - zv]

        eq3_buf = hal.dev_alloc(1 << (nv3 - 1))  # arena memory, allocated once as the prover's bump allocator would

        def run_cfg3():
            mls3 = [FoldedMultilinear(dev.slice(t << nv3, (t + 1) << nv3), 0) for t in range(5)]
            from binius_b200.layer import _u64_list

            hal._check(hal._lib.b200_tensor_product_full_query(hal._ctx, _u64_list([rr.getrandbits(128) for _ in range(nv3 - 1)]), nv3 - 1,
                                                               eq3_buf.ptr, eq3_buf.len()))
            eq3 = eq3_buf
            for r in range(nv3):
                v = nv3 - r
                be.sumcheck_compute_round_evals(v, mls3, [EqIndEvaluator(c, have_first_round_eval_1s=(r == 0)) for c in comps3], eq3, [])
                be.sumcheck_fold_multilinears(v, mls3, rr.getrandbits(128))
                if v > 1:
                    eq3 = be.fold_partial_eq_ind(v - 1, eq3)

        run_cfg3()
        hal.sync()
        ev.start()
        for _ in range(3):
            run_cfg3()
        c3_ms = ev.stop_ms() / 3
        c3_bytes = sum(16 * (5 * 2**v + 2**(v - 1)) + 16 * 5 * 2**v + 8 * 5 * 2**v for v in range(1, nv3 + 1))
        # the same 18 rounds from COMPILED host code (tools/keccak_replay.cpp cfg3, random data): per-call route and the
        # persistent kernel (csrc/tail_grid.cuh: every round in ONE cooperative kernel, challenges through host-mapped
        # mailboxes) -- the latter is the configuration's number
        comp3 = {}
        if rank == 0 and not args.no_compiled_cfg3:
            exe = os.path.join(ROOT, "tools", "keccak_replay_cpp")
            env3 = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(local_rank))
            for key, extra in (("persistent_kernel", []), ("per_call", ["notail"])):
                try:
                    hal.sync()
                    out = subprocess.run([exe, "cfg3"] + extra, capture_output=True, text=True, timeout=120, env=env3)
                    comp3[key] = json.loads(out.stdout.strip().splitlines()[-1])["ms_per_sumcheck"]
                except Exception as e:
                    comp3[key] = repr(e)
        best = comp3.get("persistent_kernel") if isinstance(comp3.get("persistent_kernel"), float) else c3_ms
        cfg3 = {"ms_per_sumcheck": best, "rounds_per_s": nv3 / (best * 1e-3), "algorithmic_bytes": c3_bytes,
                "hbm_frac": c3_bytes / (best * 1e-3) / 1e9 / peak,
                "compiled_host_persistent_kernel_ms": comp3.get("persistent_kernel"), "compiled_host_per_call_ms": comp3.get("per_call"),
                "python_mirror_per_call_ms": c3_ms,
                "note": "18 rounds incl. the eq-indicator expansion, random data; 21 MB of multilinears, so the run is latency bound "
                        "(one PCIe mailbox round trip per round + the dependent per-lane products of the small rounds), not "
                        "bandwidth bound; ms_per_sumcheck = compiled host + persistent kernel"}

    # ---- the remaining ComputeLayer ops of SURVEY.md 8a (rows a6, a7, a9, a10) at 2^22 B128 elements:
    #      device time, algorithmic GB/s and fraction of the HBM peak
    ops = None
    if not args.no_ntt:
        from binius_b200 import ArithCircuit as A, SlicesBatch, SubfieldSlice

        n22 = 1 << 22
        va, vb, vc = dev.slice(0, n22), dev.slice(n22, 2 * n22), dev.slice(2 * n22, 3 * n22)
        small = dev.slice(3 * n22, 3 * n22 + 128)
        ntt12 = binius_b200.B200AdditiveNTT(hal, 5, 24)
        expr = hal.compile_expr(A.var(0) * A.var(1) + A.var(2))
        obuf = hal.dev_alloc(2 * n22)  # outputs
        red_outs = []
        off = 0
        for r in range(22):
            red_outs.append(obuf.slice(off, off + (n22 >> (r + 1))))
            off += n22 >> (r + 1)
        chs = [rr.getrandbits(128) for _ in range(4)]
        merkle_nodes = hal.dev_alloc(2 * ((2 << 20) - 1))

        def timed(fn, alg_bytes, reps=5):
            for _ in range(2):
                hal.execute(fn)
            ev.start()
            for _ in range(reps):
                hal.execute(fn)
            ms = ev.stop_ms() / reps
            return {"ms": ms, "gbs": alg_bytes / (ms * 1e-3) / 1e9, "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / peak}

        ops = {
            "inner_product_b128_2^22": timed(lambda ex: [ex.inner_product(SubfieldSlice(va, 7), vb)], 32 * n22),
            "inner_product_b1_2^22": timed(lambda ex: [ex.inner_product(SubfieldSlice(dev.slice(0, n22 >> 7), 0), vb)], 16 * n22 + n22 // 8),
            "fold_left_b1_q128_out2^22": timed(lambda ex: (ex.fold_left(SubfieldSlice(dev.slice(0, n22), 0), small, obuf.slice(0, n22)), [])[1], 16 * n22 + 16 * n22),
            "fold_right_b1_q128_out2^22": timed(lambda ex: (ex.fold_right(SubfieldSlice(va, 0), small, obuf.slice(0, n22)), [])[1], 32 * n22),
            "compute_composite_xy+z_2^22": timed(lambda ex: (ex.compute_composite(SlicesBatch([va, vb, vc], n22), obuf.slice(0, n22), expr), [])[1], 64 * n22),
            "pairwise_product_reduce_2^22": timed(lambda ex: (ex.pairwise_product_reduce(va, red_outs), [])[1], 32 * n22),
            "merkle_commit_groestl_2^24_elems_batch16": timed(lambda ex: (hal._check(hal._lib.b200_merkle_build(hal._ctx, dev.ptr, 1 << 24, 16, merkle_nodes.ptr, (2 << 20) - 1)), [])[1], 16 * (1 << 24) + 32 * ((2 << 20) - 1), reps=3),
            "fri_fold_first_2^24_in_batch4": timed(lambda ex: (ex.fri_fold(ntt12, 20, 4, chs, dev.slice(0, 1 << 24), obuf.slice(0, 1 << 20)), [])[1], 16 * (1 << 24) + 16 * (1 << 20)),
        }

    # ---- zerocheck univariate-skip round at the keccak shape scaled to 2^24 rows:
    #      153 B1 columns, 75 degree-2 constraints, skip 7, domain 256 (the compiled replay below runs 2^27 rows)
    uni = None
    if not args.no_ntt and rank == 0:
        try:
            from binius_b200 import ArithCircuit as A
            from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals

            nvu, mu, ncu, sku = 24, 153, 75, 7  # skip 7 = the reference's choice for degree 2 over B8 (verify.rs:271-294)
            wu = 1 << (nvu - 7)
            import random as _random

            arena_u = hal.dev_alloc(mu * wu)
            hal.fill(arena_u, 0x0123456789ABCDEF0F1E2D3C4B5A6978)
            mls_u = [TransparentMultilinear(arena_u.slice(j * wu, (j + 1) * wu), 0, nvu) for j in range(mu)]
            # keccak's chi constraints out - (b0 + (b1 - 1) * b2) with their column structure (m3/src/gadgets/hash/keccak/
            # stacked.rs:318-366: 75 state_out columns, 75 b columns in three batches, round constant, 2 spare)
            comps_u = []
            for c in range(ncu):
                batch, xy = c // 25, c % 25
                x, y = xy % 5, xy // 5
                b0, b1, b2 = (A.var(75 + 25 * batch + (x + k) % 5 + 5 * y) for k in range(3))
                comps_u.append(A.var(c) - (b0 + (b1 - A.one()) * b2))
            _r = _random.Random(7)
            ch_u = [_r.getrandbits(128) for _ in range(nvu - sku)]
            be_u = B200Backend(hal)
            times = []
            for _ in range(3):
                hal.sync()
                t0 = time.perf_counter()
                o = zerocheck_univariate_evals(be_u, mls_u, comps_u, ch_u, sku, 2 << sku)  # synchronous: returns host values
                times.append((time.perf_counter() - t0) * 1e3)
                hal.dev_free(o.partial_eq_ind_evals)
            # the same round in two halves: prepare needs no challenge, finish weights the stored values
            from binius_b200.hal import zerocheck_univariate_finish, zerocheck_univariate_prepare

            halves = []
            for _ in range(3):
                hal.sync()
                t0 = time.perf_counter()
                prep_u = zerocheck_univariate_prepare(be_u, mls_u, comps_u, sku, 2 << sku)
                hal.sync()
                t1 = time.perf_counter()
                o2 = zerocheck_univariate_finish(be_u, prep_u, ch_u)
                t2 = time.perf_counter()
                halves.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
                same = o2.round_evals == o.round_evals
                hal.dev_free(o2.partial_eq_ind_evals)
                prep_u.release(be_u)
            hal.dev_free(arena_u)
            alg = mu * wu * 16 + 16 * (1 << (nvu - sku))
            uni = {"ms_per_call": min(times), "rows_log2": nvu, "columns": mu, "compositions": ncu, "skip_rounds": sku,
                   "algorithmic_bytes": alg, "hbm_frac": alg / (min(times) * 1e-3) / 1e9 / peak,
                   "two_halves": {"prepare_ms": min(h[0] for h in halves), "finish_ms": min(h[1] for h in halves), "equal_to_the_one_call_round": bool(same),
                                  "what": "b200_zerocheck_univariate_prepare (no challenge needed: overlaps with the witness upload / commitment; "
                                          "includes the allocation of the value store) + b200_zerocheck_univariate_finish"},
                   "note": "host wall time of the synchronous call (eq-ind expansion + k_uni_b8 per composition range + result copy); "
                           "shared-memory-pipe bound, see DESIGN.md section 9"}
        except Exception as e:  # never lose the headline line to the extra measurement
            uni = {"error": repr(e)}

    # ---- keccak example, n_permutations = 2^18 (BASELINE config #4): the compiled op-sequence replay
    #      (tools/keccak_replay.cpp over binius_b200/host/*.hpp) in the reference's order: witness upload + the challenge-free half
    #      of the univariate-skip round, commit (RS-encode NTT, Merkle), the other half of that round, zerocheck rounds, PIOP
    #      bivariate sumcheck (kernel scopes), FRI folds, ring-switch eq-indicators
    keccak = None
    if not args.no_ntt and rank == 0 and not args.no_keccak:
        exe = os.path.join(ROOT, "tools", "keccak_replay_cpp")
        try:
            hal.sync()
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(local_rank))
            out = subprocess.run([exe, "18"], capture_output=True, text=True, timeout=240, env=env)
            keccak = json.loads(out.stdout.strip().splitlines()[-1])
            keccak["host"] = "compiled C++ over the C ABI (binius_b200/host/*.hpp), separate process"
        except Exception as e:  # never lose the headline line to the extra measurement
            keccak = {"error": repr(e)}

    # ---- sharded sumcheck on the real layer over NCCL: low-variable partition, every round local,
    #      per-round XOR of the partial round values (all-gather), one final gather.  First a run checked round by
    #      round against the oracle (n = 16), then a timed one (n = 22).
    sharded = None
    if world > 1 and not args.no_ntt:
        import random as _rs

        from binius_b200 import sharding
        from oracle import binding as orc_chk  # checker only (never on the timed path)

        try:
            rs_ = _rs.Random(5)
            m_s, n_chk = 6, 16
            mls_chk = [orc_chk.rand_b128(100 + t, 1 << n_chk) for t in range(m_s)]
            pairs_s = [(0, 1), (2, 3), (4, 5), (1, 4)]
            al = [rs_.getrandbits(128) for _ in range(n_chk)]
            chs = [rs_.getrandbits(128) for _ in range(n_chk)]
            sc = sharding.ShardedBivariateSumcheck(hal, mls_chk, n_chk, pairs_s, world, rank, dist, comm_device=f"cuda:{local_rank}")
            cur = [x.copy() for x in mls_chk]
            exact = True
            for r in range(n_chk):
                exp = tuple(orc_chk.bivariate_round_evals(cur, n_chk - r, pairs_s, al[r]))
                exact = exact and sc.round_evals(al[r]) == exp
                sc.fold(chs[r])
                cur = [orc_chk.extrapolate_line(x[: len(x) // 2], x[len(x) // 2:], chs[r]) for x in cur]
            exact = exact and sc.finish() == [orc_chk.to_ints(x)[0] for x in cur]
            n_t = 22
            big = [np.ascontiguousarray(pinned[(t << n_t) % n_in: (t << n_t) % n_in + (1 << n_t)]) for t in range(m_s)]
            sc = sharding.ShardedBivariateSumcheck(hal, big, n_t, pairs_s, world, rank, dist, comm_device=f"cuda:{local_rank}")
            barrier()
            ts0 = time.perf_counter()
            for r in range(n_t):
                sc.round_evals(al[r % n_chk])
                sc.fold(chs[r % n_chk])
            barrier()
            sh_ms = (time.perf_counter() - ts0) * 1e3
            t_ex = torch.tensor([1 if exact else 0], device="cuda")
            dist.all_reduce(t_ex, op=dist.ReduceOp.MIN)
            sharded = {"bit_exact": bool(t_ex.item()), "checked": f"{n_chk} rounds vs the oracle on every rank (m = {m_s}, 4 pairs)",
                       "timed": {"n_vars": n_t, "multilinears": m_s, "ms": sh_ms, "rounds_per_s": n_t / (sh_ms * 1e-3),
                                 "coeffs_per_s": m_s * ((2 << n_t) - 2) / (sh_ms * 1e-3)},
                       "collective": "all_gather of 2 x B128 per round + one final gather (NCCL); no data-plane exchange"}
            # the eq-ind (zerocheck) rounds on the same partition: u32_add compositions (BASELINE config #3's), every round
            # checked against the unsharded oracle at n = 14, then config #3's own size (n = 18 after the skipped rounds)
            from binius_b200 import ArithCircuit as _A
            from binius_b200.hal import B200Backend as _Backend

            xa, ya, ci, co, zo = (_A.var(i) for i in range(5))
            comps_e = [(xa + ci) * (ya + ci) + ci - co, xa + ya + ci - zo]
            n_e = 14
            mls_e = [orc_chk.rand_b128(300 + t, 1 << n_e) for t in range(5)]
            eq_ch = [rs_.getrandbits(128) for _ in range(n_e - 1)]
            ch_e = [rs_.getrandbits(128) for _ in range(n_e)]
            be_s = _Backend(hal)
            sce = sharding.ShardedEqIndSumcheck(be_s, mls_e, n_e, comps_e, eq_ch, (), world, rank, dist, comm_device=f"cuda:{local_rank}")
            cur = [x.copy() for x in mls_e]
            eqo = orc_chk.tensor_expand(orc_chk.to_arr([1] + [0] * ((1 << (n_e - 1)) - 1)), 0, eq_ch)
            exact_e = True
            for r in range(n_e):
                vals = orc_chk.sumcheck_round_evals(1, cur, [len(x) for x in cur], [0] * 5, n_e - r, eqo, [list(c.steps) for c in comps_e],
                                                    [list(c.leading_term().steps) for c in comps_e], [1, 2], [0, 0])
                exact_e = exact_e and sce.round_evals() == [vals[0], vals[1][:1]]
                sce.fold(ch_e[r])
                cur = [orc_chk.fold_left_lerp_inplace(x, len(x), 0, n_e - r, ch_e[r]) for x in cur]
                eqo = orc_chk.fold_partial_eq_ind(eqo) if len(eqo) > 1 else eqo
            exact_e = exact_e and sce.finish() == [orc_chk.to_ints(x)[0] for x in cur]
            n_te = 18
            big_e = [np.ascontiguousarray(pinned[(t << n_te) % n_in: (t << n_te) % n_in + (1 << n_te)]) for t in range(5)]
            eq_t = [rs_.getrandbits(128) for _ in range(n_te - 1)]
            sce = sharding.ShardedEqIndSumcheck(be_s, big_e, n_te, comps_e, eq_t, (), world, rank, dist, comm_device=f"cuda:{local_rank}")
            barrier()
            ts0 = time.perf_counter()
            for r in range(n_te):
                sce.round_evals()
                sce.fold(ch_e[r % n_e])
            barrier()
            she_ms = (time.perf_counter() - ts0) * 1e3
            t_ex = torch.tensor([1 if exact_e else 0], device="cuda")
            dist.all_reduce(t_ex, op=dist.ReduceOp.MIN)
            sharded["eq_ind"] = {"bit_exact": bool(t_ex.item()), "checked": f"{n_e} rounds of the u32_add zerocheck vs the unsharded oracle on every rank",
                                 "timed": {"n_vars": n_te, "multilinears": 5, "compositions": 2, "ms": she_ms, "rounds_per_s": n_te / (she_ms * 1e-3)},
                                 "collective": "all_gather of 3 x B128 per round + one final gather (NCCL); rank g weights its share by eq(bits(g); r_low)"}
        except Exception as e:
            sharded = dict(sharded or {}, error=repr(e))

    # ---- keccak prover sharded over the ranks (BASELINE config #5: n_permutations = 2^22 over 8 GPUs):
    #      the low-variable partition gives every rank 1/N of the rows of every column, every phase of the replay runs
    #      locally on that shard (RS-encode batches, Merkle sub-trees and sumcheck rounds are independent), and the only
    #      exchange is the XOR of the round values: one all-gather of a few B128 per round, timed here on NCCL.
    #      8 ranks replay 2^19 permutations each (= 2^22); fewer ranks replay 2^18 each (config #4 per GPU, weak scaling).
    keccak_sh = None
    if world > 1 and not args.no_ntt and not args.no_keccak:
        try:
            exe = os.path.join(ROOT, "tools", "keccak_replay_cpp")
            log_shard = 19 if world == 8 else 18
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(local_rank),
                       REPLAY_PASSES="2")
            barrier()
            out = subprocess.run([exe, str(log_shard)], capture_output=True, text=True, timeout=600, env=env)
            rep = json.loads(out.stdout.strip().splitlines()[-1])
            n_rounds = 1 + (log_shard + 2) + (log_shard + 9)  # univariate round + zerocheck rounds + PIOP rounds
            buf_in = torch.zeros(4 * 75, device="cuda", dtype=torch.int64)  # up to 75 compositions x 2 values x B128 per round
            buf_out = [torch.zeros_like(buf_in) for _ in range(world)]
            for _ in range(5):
                dist.all_gather(buf_out, buf_in)
            barrier()
            tc0 = time.perf_counter()
            for _ in range(n_rounds):
                dist.all_gather(buf_out, buf_in)
                torch.cuda.synchronize()  # the host needs the values to derive the challenge
            comb_ms = (time.perf_counter() - tc0) * 1e3
            tk = torch.tensor([rep["total_ms"], -rep["total_ms"], comb_ms, rep["phases"]["witness_upload_alone"]["ms"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            t_max, t_min, comb_max, up_max = tk[0].item(), -tk[1].item(), tk[2].item(), tk[3].item()
            keccak_sh = {"n_permutations_total_log2": log_shard + int(np.log2(world)), "n_permutations_per_rank_log2": log_shard, "ranks": world,
                         "shard_replay_ms_max_over_ranks": t_max, "shard_replay_ms_min_over_ranks": t_min,
                         "round_value_combine": {"rounds": n_rounds, "ms": comb_max, "what": "all_gather of 75 x 2 B128 per round on NCCL + host sync"},
                         "total_ms": t_max + comb_max, "witness_upload_alone_ms_max_over_ranks": up_max,
                         "aggregate_witness_h2d_gbs": world * rep["phases"]["witness_upload_alone"]["h2d_bytes"] / (up_max * 1e-3) / 1e9,
                         "rank0_phases": rep["phases"],
                         "note": "every rank runs the compiled replay on its own shard concurrently (separate processes, same host); "
                                 "no data-plane collective: only the per-round values cross NVLink"}
        except Exception as e:
            keccak_sh = {"error": repr(e)}

    # ---- reduce over ranks: max time ----------------------------------------------------------------
    ms_step = ms_total / args.steps
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms, kern_ms_avg], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, kern_ms_avg = [float(x) for x in t.tolist()]
    if rank == 0:
        value = world * n_in / (ms_step * 1e-3)
        # algorithmic bytes (16 B read + 8 B written per coefficient) / average launch duration over the timed
        # region (one k_lerp_tma launch per step, CUDA events on the launching stream); the duration of
        # isolated launches (a sync between them, no overlap of one launch's tail with the next one's
        # table build) is reported next to it
        achieved = 24.0 * n_in / (ms_step * 1e-3) / 1e9
        traffic = profiled_traffic() if args.log_coeffs == 24 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u128 (GF(2^128) tower, integer/bitwise)", "data": "synthetic",
            "config": bench_config(args.log_coeffs),
            "host": {"parallelism": f"independent multilinears x{world}", "host_cpus_bound_per_rank": numa},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": world * n_in / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_in * 16, "d2h_bytes_per_step": half * 16,
                    "ms_per_step": e2e_ms},
            "roofline": {"bound": "hbm", "kernel": "k_lerp_tma", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (traffic or {}).get("bytes"), "traffic_source": (traffic or {}).get("source"),
                         "algorithmic_bytes": 24.0 * n_in, "kernel_ms": ms_step, "kernel_ms_isolated": kern_ms_avg,
                         "frac_isolated": 24.0 * n_in / (kern_ms_avg * 1e-3) / 1e9 / peak},
        }
        line["sumcheck_e2e"] = {"what": "whole sumchecks through the trait calls with ONE upload (host buffers in, final values out)",
                                "fold_chain_2^%d" % args.log_coeffs: chain}
        if biv:
            line["sumcheck_e2e"]["bivariate_m8_n20"] = biv
        if ntt_res:
            line["ntt"] = ntt_res
        if extras:
            line["sumcheck_round"] = extras
        if cfg3:
            line["zerocheck_u32_add_2^20_rows"] = cfg3
        if ops:
            line["ops"] = ops
        if uni:
            line["zerocheck_univariate_skip_2^24_rows"] = uni
        if keccak:
            line["keccak_replay_2^18"] = keccak
        if sharded:
            line["sharded_sumcheck"] = sharded
        if keccak_sh:
            line["keccak_replay_sharded"] = keccak_sh
        if world > 1:
            line["e2e"]["aggregate_h2d_gbs"] = world * n_in * 16 / (e2e_ms * 1e-3) / 1e9
        if not args.no_cpu and world == 1:
            os.sched_setaffinity(0, all_cpus)  # the CPU arms use every host core
            line["cpu_baseline"] = cpu_fold_baseline(args.log_coeffs)
            from oracle import binding as orc

            def arm(fn):
                try:
                    return fn()
                except Exception as e:
                    return {"error": repr(e)}

            chain["cpu_baseline"] = arm(lambda: orc.cpu_fold_chain_parallel(args.log_coeffs, 5.0))
            if biv:
                biv["cpu_baseline"] = arm(lambda: orc.cpu_bivariate_sumcheck_parallel(8, 20, 8, 5.0))
            if cfg3:
                cfg3["cpu_baseline"] = arm(lambda: orc.cpu_u32add_zerocheck_parallel(18, 4.0))
            if extras and "tensor_expand_k22" in extras:
                extras["tensor_expand_k22"]["cpu_baseline"] = arm(lambda: orc.cpu_tensor_expand_parallel(22, 3.0))
            if ntt_res:
                ntt_res["S1_rs_encode"]["cpu_baseline"] = arm(lambda: orc.cpu_ntt_parallel(6, 18, 1, 5.0, d=24))
            if uni and "error" not in uni:
                uni["value"] = uni["columns"] * (1 << (uni["rows_log2"] - uni["skip_rounds"])) / (uni["ms_per_call"] * 1e-3)
                uni["unit"] = "sub-cube columns/s"
                uni["cpu_baseline"] = arm(cpu_univariate_baseline)
            if keccak and "error" not in keccak:
                keccak["cpu_baseline"] = arm(cpu_keccak_replay_baseline)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    hal.close()


if __name__ == "__main__":
    main()
