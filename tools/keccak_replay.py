#!/usr/bin/env python3
"""Kernel-level replay of the hot-path op sequence of the reference's keccak example prover
(examples/keccak.rs, n_permutations = 2^log_n) on synthetic data -- BASELINE configs #4/#5.

The Rust prover cannot be linked here (no toolchain), so this drives the SAME ops with the SAME shapes
through the C ABI and reports device time per phase (SURVEY.md 8d "cfg#4/#5 ... otherwise report the
kernel-level replay of the same op sequence and say so").  Shapes (keccak
gadget m3/src/gadgets/hash/keccak/stacked.rs):
  commit      RS-encode NTT over B32 of the 2^(log_n+10) B128 codeword: log_x = 6, log_y = log_n+6, skip 1
  zerocheck   153 multilinears of 2^(log_n+2) B128 (after the 7-round univariate skip), 75 degree-2 chi
              constraints out - (b0 + (b1-1)*b2), eq-ind 2^(log_n+1): log_n+2 rounds of
              {eq-ind round evals at 1 and infinity, fold all multilinears, halve eq-ind}
  piop        bivariate-product sumcheck over 200 multilinears (100 committed x 100 transparent) of
              2^(log_n+2), 100 pairs: log_n+2 rounds of {round evals, fold all}
  fri         first fold of the codeword with log_batch = 4, then arity-4 folds down to 2^12
  ring_switch 8 x tensor_expand (k = log_n+2) + fold_right (kappa = 7 row-batch coefficients)
  univariate  (reported apart, not in total_ms) the zerocheck univariate-skip round on 153 B1 columns of
              2^(log_n+9) rows, 75 degree-2 constraints, skip 6 / domain 128 (the shape the fast kernel covers
              in one launch; the reference's skip for this constraint degree is 7)
What is NOT replayed (out of scope, stays on the host in the reference): witness generation,
GKR grand product / exponentiation, evalcheck, Merkle hashing, transcript.

NOTE: the figures the bench line and the documents report come from the COMPILED replay, tools/keccak_replay.cpp, which
has since grown the phases this Python version lacks -- witness upload with the univariate-skip round prepared behind it
(skip 7), Merkle commit, univariate finish, the projection onto the remaining rounds -- and runs them in the reference's
order.  This file remains as the Python-mirror view of the sumcheck / NTT / FRI phases (host overhead of the mirror).
"""
import argparse
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import binius_b200
from binius_b200 import ArithCircuit as A
from binius_b200 import NTTShape, SubfieldSlice
from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear


class Timer:
    def __init__(self, hal):
        self.hal = hal
        self.a, self.b = C.c_void_p(), C.c_void_p()
        hal._check(hal._lib.b200_event_create(hal._ctx, C.byref(self.a)))
        hal._check(hal._lib.b200_event_create(hal._ctx, C.byref(self.b)))

    def __enter__(self):
        self.hal.sync()
        self.l0 = self.hal.launch_count()
        self.hal._check(self.hal._lib.b200_event_record(self.hal._ctx, self.a))
        return self

    def __exit__(self, *exc):
        self.hal._check(self.hal._lib.b200_event_record(self.hal._ctx, self.b))
        ms = C.c_float()
        self.hal._check(self.hal._lib.b200_event_elapsed_ms(self.hal._ctx, self.a, self.b, C.byref(ms)))
        self.ms = float(ms.value)
        self.launches = self.hal.launch_count() - self.l0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=18, help="log2 of the number of keccak-f permutations")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--warm", type=int, default=1, help="1: run the sequence once untimed first")
    args = ap.parse_args()
    rng = random.Random(0)
    hal = binius_b200.B200Layer(args.device)
    be = B200Backend(hal)
    nv = args.log_n + 2  # variables of a folded multilinear
    res = {"workload": f"keccak op-sequence replay, n_permutations = 2^{args.log_n} (synthetic data)", "phases": {}}

    # one big arena filled with a non-trivial pattern (content does not affect the work done)
    n_code = 1 << (args.log_n + 10)
    arena_elems = max(n_code, 200 << nv)
    arena = hal.dev_alloc(arena_elems)
    hal.fill(arena, 0x0123456789ABCDEF_FEDCBA9876543211)

    # pass 0 warms the context (scratch pool, lazily loaded kernels) as a long-lived prover would be; pass 1 is reported
    for pass_ix in range(2 if args.warm else 1):
        # ---- commit: RS encode NTT
        ntt = binius_b200.B200AdditiveNTT(hal, 5, args.log_n + 6)
        with Timer(hal) as t:
            ntt.forward_device(arena.ptr, 5, n_code * 4, NTTShape(6, args.log_n + 6, 0), 0, 0, 1)
        res["phases"]["commit_rs_encode_ntt"] = {"ms": t.ms, "launches": t.launches, "coeffs_b32": n_code * 4}

        # ---- zerocheck multilinear rounds
        m = 153
        mls = [FoldedMultilinear(arena.slice(i << nv, (i + 1) << nv), 0) for i in range(m)]
        comps = []
        for c in range(75):
            out, b0, b1, b2 = (A.var(k) for k in (c, 75 + (c % 26), 75 + ((c + 1) % 26), 75 + ((c + 2) % 26)))
            comps.append(out - (b0 + (b1 - A.one()) * b2))
        eq = be.tensor_product_full_query([rng.getrandbits(128) for _ in range(nv - 1)])
        t_ev = t_fold = 0.0
        launches = 0
        for rnd in range(nv):
            v = nv - rnd
            evs = [EqIndEvaluator(c, have_first_round_eval_1s=(rnd == 0)) for c in comps]
            with Timer(hal) as t:
                be.sumcheck_compute_round_evals(v, mls, evs, eq, [])
            t_ev += t.ms
            launches += t.launches
            if os.environ.get("REPLAY_VERBOSE"):
                print(f"zerocheck round n_vars={v}: evals {t.ms:.3f} ms, {t.launches} launches", file=sys.stderr)
            with Timer(hal) as t:
                be.sumcheck_fold_multilinears(v, mls, rng.getrandbits(128))
                if v > 1:
                    eq = be.fold_partial_eq_ind(v - 1, eq)
            t_fold += t.ms
            launches += t.launches
            if os.environ.get("REPLAY_VERBOSE"):
                print(f"zerocheck round n_vars={v}: fold {t.ms:.3f} ms, {t.launches} launches", file=sys.stderr)
        res["phases"]["zerocheck_rounds"] = {"round_evals_ms": t_ev, "fold_ms": t_fold, "ms": t_ev + t_fold, "launches": launches,
                                             "multilinears": m, "compositions": 75, "n_vars": nv}

        # ---- PIOP bivariate sumcheck
        hal.fill(arena, 0x0F1E2D3C4B5A6978_8796A5B4C3D2E1F1)
        m2 = 200
        pml = [FoldedMultilinear(arena.slice(i << nv, (i + 1) << nv), 0) for i in range(m2)]
        pairs = [(i, 100 + i) for i in range(100)]
        t_ev = t_fold = 0.0
        launches = 0
        for rnd in range(nv):
            v = nv - rnd
            alpha = rng.getrandbits(128)
            dml = [x.evals for x in pml]
            with Timer(hal) as t:
                hal.execute(lambda ex: list(ex.bivariate_round_evals(dml, v, pairs, alpha)))
            t_ev += t.ms
            launches += t.launches
            if os.environ.get("REPLAY_VERBOSE"):
                print(f"piop round n_vars={v}: evals {t.ms:.3f} ms, {t.launches} launches", file=sys.stderr)
            # the prover issues one extrapolate_line per multilinear inside one execute (v3/bivariate_product.rs:
            # 196-232); the library batches them into multi-segment launches.  One C-ABI call here keeps the
            # python call overhead (which a Rust host would not have) out of the device timing.
            with Timer(hal) as t:
                be.sumcheck_fold_multilinears(v, pml, rng.getrandbits(128))
            t_fold += t.ms
            launches += t.launches
        res["phases"]["piop_bivariate_sumcheck"] = {"round_evals_ms": t_ev, "fold_ms": t_fold, "ms": t_ev + t_fold, "launches": launches,
                                                    "multilinears": m2, "compositions": 100, "n_vars": nv}

        # ---- FRI folds
        fri_ntt = binius_b200.B200AdditiveNTT(hal, 5, min(32, args.log_n + 10))
        cur_log = args.log_n + 6  # log_len after removing the log_batch = 4 interleave
        out = hal.dev_alloc(1 << cur_log)
        t_fri = 0.0
        launches = 0
        with Timer(hal) as t:
            hal.execute(lambda ex: (ex.fri_fold(fri_ntt, cur_log, 4, [rng.getrandbits(128) for _ in range(4)], arena.slice(0, n_code), out), [])[1])
        t_fri += t.ms
        launches += t.launches
        src = out
        while cur_log - 4 >= 12:
            dst = hal.dev_alloc(1 << (cur_log - 4))
            ch = [rng.getrandbits(128) for _ in range(4)]
            with Timer(hal) as t:
                hal.execute(lambda ex: (ex.fri_fold(fri_ntt, cur_log, 0, ch, src, dst), [])[1])
            t_fri += t.ms
            launches += t.launches
            src, cur_log = dst, cur_log - 4
        res["phases"]["fri_folds"] = {"ms": t_fri, "launches": launches, "first_fold_inputs": n_code}

        # ---- ring switch eq-inds
        t_rs = 0.0
        launches = 0
        q = hal.to_device(binius_b200.to_arr([rng.getrandbits(128) for _ in range(128)]))
        for claim in range(8):
            buf = arena.slice(claim << nv, (claim + 1) << nv)
            hal.fill(buf.slice(0, 1), 1)
            coords = [rng.getrandbits(128) for _ in range(nv)]
            mle = hal.dev_alloc(1 << nv)
            with Timer(hal) as t:
                hal.execute(lambda ex: (ex.tensor_expand(0, coords, buf), ex.fold_right(SubfieldSlice(buf, 0), q, mle), [])[2])
            t_rs += t.ms
            launches += t.launches
            hal.dev_free(mle)
        res["phases"]["ring_switch_eq_inds"] = {"ms": t_rs, "launches": launches, "claims": 8}

    # ---- zerocheck univariate-skip round (host wall time of the synchronous call; reported apart)
    try:
        import time

        from binius_b200.hal import TransparentMultilinear, zerocheck_univariate_evals

        n_rows = args.log_n + 9
        words = 1 << (n_rows - 7)
        cols = [TransparentMultilinear(arena.slice(j * words, (j + 1) * words), 0, n_rows) for j in range(153)]
        ucomps = [A.var((2 * c) % 153) * A.var((2 * c + 1) % 153) + A.var((2 * c + 5) % 153) + A.var((2 * c + 11) % 153) for c in range(75)]
        chs = [rng.getrandbits(128) for _ in range(n_rows - 6)]
        best = None
        for _ in range(2):
            hal.sync()
            t0 = time.perf_counter()
            o = zerocheck_univariate_evals(be, cols, ucomps, chs, 6, 128)
            dt = (time.perf_counter() - t0) * 1e3
            hal.dev_free(o.partial_eq_ind_evals)
            best = dt if best is None else min(best, dt)
        res["univariate_skip_round"] = {"ms": best, "rows_log2": n_rows, "columns": 153, "compositions": 75, "skip_rounds": 6}
    except Exception as e:  # the replay proper must not depend on this extra
        res["univariate_skip_round"] = {"error": repr(e)}

    res["warm_pass"] = bool(args.warm)
    res["total_ms"] = sum(p["ms"] for p in res["phases"].values())
    res["gpu_launches"] = sum(p["launches"] for p in res["phases"].values())
    print(json.dumps(res))
    hal.close()


if __name__ == "__main__":
    main()
