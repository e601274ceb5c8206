#!/usr/bin/env python
"""bench.py -- the simplex tableau pivot loop on B200 (BASELINE.json metric: pivots/s and pivot-update HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pivots P]

A "step" is one pass of the hot path over one batch of synthetic input: P pivots of the device-resident loop.
  N = 1 : BASELINE config 4 -- one dense 16384 x 16384 fp64 tableau (2.1 GB > L2), Bland rule, fixed budget.
  N > 1 : BASELINE config 5 -- one dense 131072 x 131072 fp64 tableau (137 GB) column-sharded over the N GPUs,
          one all-gather of candidate columns per pivot (launched by torchrun, one rank per GPU).
`value`   = pivot-update GB/s of the whole job (pivots/s x 2*R*C*8 bytes, all GPUs) with the tableau already resident
            in HBM (max over ranks, CUDA events); `pivots_per_s` is printed beside it.
`e2e`     = the same metric through the reference-facing C-ABI call with HOST buffers (b200lp_solve_dense from
            pinned host arrays A, b, c -> x, z), host<->device copies inside the timed region.
`roofline`= pivot-update kernel: algorithmic bytes per launch (2*R*C*8) / its average duration, measured live with
            CUDA events around every launch of a real loop; peak = MEASURED_PEAKS.json hbm_gbs (else the fallback).
`cpu_baseline` = oracle/ (the CPU restatement, OpenMP) timed on this box's host cores on a bounded sample.
`--impl reference` times that CPU path alone as the reference arm (the reference itself holds no pivot loop:
its engine is the un-vendored simple-simplex package).
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
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md
# BASELINE.json's metric is "pivots/sec and pivot-update HBM GB/s".  `value` is the GB/s form -- pivots/s x 2*R*C*8
# bytes, summed over the GPUs -- because it is comparable between the 1-GPU tableau (config 4) and the sharded one
# (config 5) that the 1/2/4/8-GPU series is made of; pivots/s is printed beside it as `pivots_per_s`.
METRIC = "pivot_update_hbm_GBps"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if os.environ.get("B200LP_BENCH_NO_SAMPLER"):  # diagnostic: is the 50 ms nvidia-smi poll visible in the timing?
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def wait_first(self, timeout=5.0):
        t0 = time.monotonic()
        while not self.lines and time.monotonic() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        """Only samples that arrive after this call are reported (call it when the timed region starts)."""
        self.t_mark = time.monotonic()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, ln in self.lines:
            if ts < t_mark:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# FP64 FMA issue rate of one B200 measured on this pool with a pure DFMA kernel (scripts/probe_dmma.cu,
# profiles/probe_dmma_r2.txt: 15.97-16.22 T FMA/s at the clock the power cap allows, ~1.69 GHz x 148 SMs x 64 FMA/clk)
FP64_FMA_PER_S_MEASURED = 16.0e12


def lookahead_roofline(pps, R, C, K, gpus=1):
    """The look-ahead loop against its two roofs: the flush moves the tableau once per K pivots (2*R*C*8 bytes) and issues
    one FP64 FMA per element and pivot (R*C per pivot), per GPU."""
    peak, _ = measured_peak()
    hbm = pps * 16.0 * R * C / K / gpus / 1e9
    fma = pps * float(R) * float(C) / gpus
    return {"hbm_GBps_per_gpu": hbm, "frac_of_hbm_peak": hbm / peak, "fp64_fma_per_s_per_gpu": fma,
            "frac_of_fp64_peak": fma / FP64_FMA_PER_S_MEASURED,
            "fp64_peak_source": "measured DFMA issue rate, 16.0e12 FMA/s (scripts/probe_dmma.cu)"}


def nvlink_counters(gpu_index):
    """Sum of the NVLink data counters of one GPU in bytes (tx, rx, source), or None.  NVML field values first
    (NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX, KiB, summed over the links by the driver), else `nvidia-smi nvlink
    -gt d` ("Data Tx: N KiB" per link).  Read before and after the timed region of the sharded bench: the difference is
    what the peer-memory exchange moved."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX,
                                                   pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                raise RuntimeError(f"field {v.fieldId}: nvmlReturn {v.nvmlReturn}")
            out.append(int(v.value.ullVal))
        return 1024 * out[0], 1024 * out[1], "NVML field values NVLINK_THROUGHPUT_DATA_TX/RX (KiB, all links)"
    except Exception:  # noqa: BLE001
        pass
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu_index)], capture_output=True, text=True,
                             timeout=20).stdout
    except Exception:
        return None
    import re
    tx = [int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    rx = [int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
    if not tx and not rx:
        return None
    return 1024 * sum(tx), 1024 * sum(rx), "nvidia-smi nvlink -gt d, all links"


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_pivots_per_s(R, C_total, rule, pivots, seed, threads=None, repeats=1):
    """Oracle (OpenMP) on a generated R x C_total tableau: (pivots/s, threads, seconds)."""
    from oracle import oracle as O
    O.build()
    threads = threads or O.host_cores()
    O.set_threads(threads)
    t = O.OracleTableau.generate(seed, R - 1, C_total - 1)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = t.solve(O.make_opts(rule=rule, max_pivots=pivots, threads=threads))
        dt = time.perf_counter() - t0
        pps = res["n_pivots"] / dt
        best = pps if best is None else max(best, pps)
    return best, threads, dt


def run_reference_arm(args):
    """--impl reference: the CPU implementation of the path on this box's host cores, same config/metric/unit.

    Threads = every core this process may run on (its affinity mask), passed to OpenMP explicitly: torchrun exports
    OMP_NUM_THREADS=1 to its workers, which is a launcher default, not a property of the CPU path.  A 1-core figure of
    the same sample is printed beside the all-core one ("1 core and all host cores")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R = args.rows
    C = args.cols_total
    extrapolated = False
    if args.gpus > 1:
        # config 5 (137 GB) does not fit host-side timing budgets: time a 16384-column slab of it and scale by columns
        C = 16384
        extrapolated = True
    from oracle import oracle as O
    O.build()
    threads = O.host_cores()
    O.set_threads(threads)
    rule = 1 if args.rule == "bland" else 0
    tab = O.OracleTableau.generate(args.seed, R - 1, args.cols_total - 1, 0, C - 1)
    opts = O.make_opts(rule=rule, max_pivots=args.ref_pivots, threads=threads)
    for _ in range(args.warmup):
        tab.solve(opts)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += tab.solve(opts)["n_pivots"]
    dt = time.perf_counter() - t0
    scale = C / args.cols_total
    pps = n / dt * scale
    # the same sample on ONE core (a quarter of the pivots of one step: it is ~`threads` times slower)
    one_piv = max(2, args.ref_pivots // 4)
    t1 = time.perf_counter()
    n1 = tab.solve(O.make_opts(rule=rule, max_pivots=one_piv, threads=1))["n_pivots"]
    pps1 = n1 / (time.perf_counter() - t1) * scale
    bpp = 16.0 * args.rows * args.cols_total
    gbps = pps * bpp / 1e9
    note = (f"{args.ref_pivots} pivots per step on " +
            (f"a {R} x {C} column slab of the {R} x {args.cols_total} tableau, scaled by {C}/{args.cols_total} (extrapolated)"
             if extrapolated else f"the full {R} x {C} tableau") +
            f"; oracle/ (C restatement, OpenMP) on {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": "GB/s", "pivots_per_s": pps, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "reference_sample": {"pivots_per_step": args.ref_pivots, "rows": R, "cols": C, "extrapolated": extrapolated,
                             "scale_to_full_tableau": scale, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": gbps, "unit": "GB/s", "pivots_per_s": pps, "cores": threads, "kind": "port",
                         "sample": note, "host_cpus": os.cpu_count(),
                         "one_core": {"value": pps1 * bpp / 1e9, "unit": "GB/s", "pivots_per_s": pps1, "cores": 1,
                                      "sample": f"{one_piv} pivots of the same sample"}},
        "e2e": {"value": gbps, "unit": "GB/s", "pivots_per_s": pps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    if args.gpus == 1:
        return {"workload": f"BASELINE config 4: single dense {args.rows}x{args.rows} fp64 condensed tableau, "
                            f"{args.rule} rule, fixed budget of {args.pivots} pivots per step",
                "rows": args.rows, "cols": args.rows, "rule": args.rule, "pivots_per_step": args.pivots,
                "generator": getattr(args, "generator", "counter"),
                "bytes_per_pivot": 16 * args.rows * args.rows,
                "l2": f"tableau ({8 * args.rows * args.rows / 1e9:.2f} GB) vs L2 (0.126 GB): every pivot streams it from HBM",
                "parallelism": "1 GPU"}
    return {"workload": f"BASELINE config 5: single dense {args.rows}x{args.cols_total} fp64 condensed tableau "
                        f"column-sharded over {args.gpus} GPUs, {args.rule} rule, {args.pivots} pivots per step",
            "rows": args.rows, "cols": args.cols_total, "rule": args.rule, "pivots_per_step": args.pivots,
            "bytes_per_pivot": 16 * args.rows * args.cols_total,
            "l2": "each shard is far larger than L2", "parallelism": f"column-sharded x{args.gpus}, "
            "1 exchange of candidate columns per pivot (see 'collective')"}


# ----------------------------------------------------------------------------------------------------------
# parity of the timed run: pivot history against the committed oracle sequence
# ----------------------------------------------------------------------------------------------------------
def expected_history(args):
    """(row, entering id, leaving id) of the first pivots of this config under Bland, computed by the CPU oracle
    (tests/golden/make_pivot_history.py).  None when the run is not one of the two committed configs."""
    if args.rule != "bland" or args.seed != 4 or getattr(args, "generator", "counter") != "counter":
        return None
    name = {(16384, 16384): "config4", (131072, 131072): "config5"}.get((args.rows, args.cols_total))
    if not name:
        return None
    try:
        return np.load(os.path.join(ROOT, "tests", "golden", f"pivot_history_{name}.npy"))
    except Exception:
        return None


class HistoryCheck:
    """Collects the pivot history of consecutive steps that start from a freshly generated tableau and compares it with
    the oracle's sequence; the SHA-256 is over the int32 triples (row, entering id, leaving id) in pivot order."""

    def __init__(self, args):
        self.want = expected_history(args)
        self.rows = []

    def add(self, h):
        self.rows.append(np.stack([h["piv_row"], h["enter_lab"], h["leave_lab"]], axis=1).astype(np.int32))

    def report(self):
        import hashlib
        got = np.concatenate(self.rows, axis=0) if self.rows else np.zeros((0, 3), np.int32)
        out = {"pivots": int(len(got)), "sha256": hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest()}
        if self.want is not None:
            k = min(len(got), len(self.want))
            out["compared_with_oracle"] = int(k)
            out["equals_oracle"] = bool(np.array_equal(got[:k], self.want[:k]))
            out["oracle_sha256"] = hashlib.sha256(np.ascontiguousarray(self.want[:len(got)]).tobytes()).hexdigest() \
                if len(self.want) >= len(got) else None
        return out


def single_gpu_parity_cases(device=0):
    """Small invocations of every loop of the single-GPU path, bit for bit against the oracle (pivots and tableau)."""
    from oracle import oracle as O
    from simplex_solver_b200 import native
    import torch
    s = native.Solver(device)
    cases, ok = 0, True
    for (m, n, budget) in ((511, 767, 96), (130, 257, 60)):
        ld = (n + 1 + 15) // 16 * 16
        for rule in (native.RULE_BLAND, native.RULE_DANTZIG):
            ot = O.OracleTableau.generate(4, m, n)
            ref = ot.solve(O.make_opts(rule=rule, max_pivots=budget), hist_cap=budget)
            for kw in (dict(loop_mode=native.LOOP_GRAPH, update_variant=native.UPDATE_LDG),
                       dict(loop_mode=native.LOOP_GRAPH, update_variant=native.UPDATE_TMA),
                       dict(loop_mode=native.LOOP_AUTO), dict(loop_mode=native.LOOP_BLOCKED, check_every=24)):
                T = torch.empty((m + 1) * ld, dtype=torch.float64, device=f"cuda:{device}")
                s.attach(T.data_ptr(), m, 1, n + 1, ld, n, n + m, keep=T)
                s.generate(4, n, 0)
                got = s.run(native.make_opts(rule=rule, max_pivots=budget, **kw), hist_cap=budget)
                good = (got["n_pivots"] == ref["n_pivots"] and np.array_equal(got["piv_row"], ref["piv_row"])
                        and np.array_equal(got["enter_lab"], ref["enter_lab"]) and np.array_equal(s.read_tableau(), ot.T))
                cases += 1
                ok = ok and bool(good)
    s.close()
    return cases, ok


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def bench_single_gpu(args):
    import torch
    torch.cuda.set_device(0)
    # one explicit non-default stream carries the solver's kernels, torch's copies and the timing events
    with torch.cuda.stream(torch.cuda.Stream()):
        _bench_single_gpu(args)


def _bench_single_gpu(args):
    import torch
    from simplex_solver_b200 import native

    R = args.rows
    m = n = R - 1
    C = ld = R
    rule = native.RULE_BLAND if args.rule == "bland" else native.RULE_DANTZIG
    variant = {"auto": native.UPDATE_AUTO, "ldg": native.UPDATE_LDG, "tma": native.UPDATE_TMA}[args.variant]
    s = native.Solver(0)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    T = torch.empty(R * ld, dtype=torch.float64, device="cuda:0")
    s.attach(T.data_ptr(), m, 1, C, ld, n, n + m, keep=T)
    if args.generator == "survey":
        # SURVEY.md 8d's literal definition of config 4: the config-2 generator (numpy default_rng) with seed 4,
        # A ~ U[0,1), b = A x0 + U[0.1,1), c ~ U[0.1,1), maximise -- built on the host once and uploaded per refresh
        from simplex_solver_b200 import workloads as W
        A_s, b_s, c_s, ops_s, _ = W.dense_feasible_lp(n, seed=args.seed, m=m)
        T.view(R, ld)[:m, :n].copy_(torch.from_numpy(A_s))
        T.view(R, ld)[:m, n].copy_(torch.from_numpy(b_s))
        T.view(R, ld)[m, :n].copy_(torch.from_numpy(-c_s))
        T.view(R, ld)[m, n] = 0.0
        torch.cuda.synchronize()
        T0 = T.clone()
        rl0 = np.concatenate([n + np.arange(m), [-1]]).astype(np.int32)
        cl0 = np.concatenate([np.arange(n), [-1]]).astype(np.int32)
        del A_s

        def fresh():
            s.attach(T.data_ptr(), m, 1, C, ld, n, n + m, keep=T)   # also clears the loop state
            T.copy_(T0)
            s.set_labels(rl0, cl0)
    else:
        def fresh():
            s.generate(args.seed, n, 0)
    fresh()
    torch.cuda.synchronize()
    # the headline is the rank-1 loop that north_star specifies (one fused tableau pass per pivot): LOOP_GRAPH is set
    # explicitly because the library's AUTO mode would pick the look-ahead loop for a tableau of this size
    opts = native.make_opts(rule=rule, max_pivots=args.pivots, update_variant=variant, loop_mode=native.LOOP_GRAPH)
    bytes_per_pivot = 16.0 * R * C

    # parity before anything is timed: every loop of the path on small tableaux, bit for bit against the oracle
    parity = None
    if args.parity:
        n_cases, ok = single_gpu_parity_cases(0)
        parity = {"cases": n_cases, "ok": ok, "checker": "oracle/ (pivot history and tableau bits)"}

    launches = 0
    sampler = ClockSampler(0)
    sampler.start()
    sampler.wait_first()
    for _ in range(args.warmup):
        s.run(opts)
    # the timed steps start from the freshly generated tableau, so that their pivot history is a prefix of the sequence
    # the oracle computed for this config (tests/golden/pivot_history_config4.npy)
    fresh()
    torch.cuda.synchronize()
    hist = HistoryCheck(args)
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pivots = 0
    for _ in range(args.steps):
        r = s.run(opts, hist_cap=args.pivots)
        pivots += r["n_pivots"]
        launches += r["kernel_launches"]
        hist.add(r)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    sec = ev0.elapsed_time(ev1) * 1e-3
    value = pivots / sec
    if parity is not None:
        parity["timed_history"] = hist.report()
        parity["ok"] = parity["ok"] and parity["timed_history"].get("equals_oracle", True)

    # live per-kernel durations of a real loop (events around every launch)
    prof = s.profile_loop(opts, iters=min(64, args.pivots))
    peak, peak_src = measured_peak()
    achieved = bytes_per_pivot / (prof["update_ms"] * 1e-3) / 1e9
    ncu_traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "update_kernel_traffic.json")) as f:
            ncu_traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_update_ldg<256,8>" if variant != native.UPDATE_TMA else "k_update_tma",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": ncu_traffic,
                "algorithmic_bytes_per_launch": bytes_per_pivot,
                "kernel_ms": {"update": prof["update_ms"], "price": prof["price_ms"], "ratio": prof["ratio_ms"]},
                "update_share_of_pivot": prof["update_ms"] / (prof["update_ms"] + prof["price_ms"] + prof["ratio_ms"]),
                "loop_GBps": value * bytes_per_pivot / 1e9, "loop_frac_of_peak": value * bytes_per_pivot / 1e9 / peak,
                "loop_frac_of_nominal_8TBs": value * bytes_per_pivot / 1e9 / 8000.0}

    # ---- beyond the rank-1 roofline: look-ahead pivoting (same pivots, same tableau bits, 1 tableau pass per K) ----
    lookahead = None
    if args.lookahead:
        lookahead = {}
        for K in (8, 16, 32):
            fresh()
            ob = native.make_opts(rule=rule, max_pivots=args.pivots * 4, loop_mode=native.LOOP_BLOCKED, check_every=K)
            s.run(ob)
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fresh()
            torch.cuda.synchronize()
            b0.record()
            rb = s.run(ob, hist_cap=args.pivots * 4)
            b1.record()
            torch.cuda.synchronize()
            pps = rb["n_pivots"] / (b0.elapsed_time(b1) * 1e-3)
            hk = HistoryCheck(args)
            hk.add(rb)
            lookahead[f"K={K}"] = {"pivots_per_s": pps, "speedup_vs_rank1_loop": pps / value,
                                   "hbm_bytes_per_pivot": bytes_per_pivot / K, "kernel_launches": rb["kernel_launches"],
                                   "roofline": lookahead_roofline(pps, R, C, K), "history": hk.report()}
            if parity is not None:
                parity["ok"] = parity["ok"] and lookahead[f"K={K}"]["history"].get("equals_oracle", True)
        lookahead["note"] = ("loop_mode=BLOCKED: K pivots are decided from O(R+C) state and applied in one pass over the "
                             "tableau; pivot sequence and tableau are bit-identical to the rank-1 loop (tests), HBM traffic "
                             "per pivot is 2*R*C*8/K, so pivots/s is no longer bounded by the rank-1 roofline")

    # ---- e2e: the reference-facing call with HOST buffers (pinned), copies inside the timed region ----
    fresh()
    torch.cuda.synchronize()
    Th = torch.empty((R, ld), dtype=torch.float64, pin_memory=True)
    Th.copy_(T.view(R, ld))
    torch.cuda.synchronize()
    A_h = Th[:m, :n]                      # row stride ld: passed as lda
    b_h = Th[:m, n].clone().pin_memory()
    c_h = Th[m, :n].clone().pin_memory()  # objective row = costs of the minimisation form
    ops_h = np.zeros(m, dtype=np.int8)
    s2 = native.Solver(0)
    s2.set_stream(torch.cuda.current_stream().cuda_stream)
    del T
    s.close()
    torch.cuda.empty_cache()

    def e2e_step():
        return _solve_dense_host_ptr(s2, A_h, b_h, c_h, ops_h, opts, m, n, ld)

    e2e_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(1, min(args.steps, 3))
    e0.record()
    piv2 = 0
    for _ in range(e2e_steps):
        r = e2e_step()
        piv2 += r["n_pivots"]
    e1.record()
    torch.cuda.synchronize()
    e2e_sec = e0.elapsed_time(e1) * 1e-3
    e2e = {"value": piv2 / e2e_sec * bytes_per_pivot / 1e9, "unit": "GB/s", "pivots_per_s": piv2 / e2e_sec,
           "h2d_bytes_per_step": int(8 * (m * n + m + n)),
           "d2h_bytes_per_step": int(8 * (n + 1) + 128), "steps": e2e_steps, "ms_per_step": e2e_sec / e2e_steps * 1e3,
           "call": "b200lp_solve_dense(A, b, c, ops from pinned host memory) -> x, c'x on the host"}
    if lookahead is not None:
        # the same host-buffer call with the library's default loop_mode (AUTO -> look-ahead loop at this size)
        o_auto = native.make_opts(rule=opts.rule, max_pivots=args.pivots)
        _solve_dense_host_ptr(s2, A_h, b_h, c_h, ops_h, o_auto, m, n, ld)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        pa = 0
        for _ in range(e2e_steps):
            pa += _solve_dense_host_ptr(s2, A_h, b_h, c_h, ops_h, o_auto, m, n, ld)["n_pivots"]
        a1.record()
        torch.cuda.synchronize()
        lookahead["e2e_default_loop_mode"] = {
            "pivots_per_s": pa / (a0.elapsed_time(a1) * 1e-3), "ms_per_step": a0.elapsed_time(a1) / e2e_steps,
            "call": "b200lp_solve_dense from pinned host memory with loop_mode = AUTO (the default)"}
    s2.close()
    del Th
    torch.cuda.empty_cache()

    # ---- CPU baseline on the host cores (bounded sample) ----
    cpu = None
    if not args.no_cpu:
        cpu_piv = args.cpu_pivots
        pps, threads, dt = cpu_pivots_per_s(R, C, 1 if args.rule == "bland" else 0, cpu_piv, args.seed)
        cpu = {"value": pps * bytes_per_pivot / 1e9, "unit": "GB/s", "pivots_per_s": pps, "cores": threads, "kind": "port",
               "host_cpus": os.cpu_count(),
               "sample": f"{cpu_piv} pivots of the same {R}x{C} tableau with oracle/ (OpenMP), {dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value * bytes_per_pivot / 1e9, "unit": "GB/s", "pivots_per_s": value, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if parity is not None:
        line["parity"] = parity
    if lookahead:
        line["lookahead"] = lookahead
    if args.secondary:
        line["secondary"] = secondary_configs(args)
    if args.config5_one_gpu:
        line["config5_one_gpu"] = config5_on_one_gpu(args)
    print(json.dumps(line))
    if parity is not None and not parity["ok"]:
        print("[bench] PARITY FAILURE: the GPU path differs from the oracle", file=sys.stderr)
        sys.exit(3)


def config5_on_one_gpu(args):
    """BASELINE config 5's tableau (131072 x 131072 fp64, 137.4 GB) on ONE B200: the same LP the 2/4/8-GPU lines
    shard, as a same-workload anchor of the 1 -> 8 curve (non-headline).  Skipped when the device has no room."""
    import torch
    from simplex_solver_b200 import native
    R = C = 131072
    need = 8 * R * C + (6 << 30)
    torch.cuda.empty_cache()
    free, total = torch.cuda.mem_get_info()
    if free < need:
        return {"skipped": f"needs {need / 1e9:.1f} GB of HBM, {free / 1e9:.1f} GB free"}
    try:
        T = torch.empty(R * C, dtype=torch.float64, device="cuda:0")
        s = native.Solver(0)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        m = n = R - 1
        s.attach(T.data_ptr(), m, 1, C, C, n, n + m, keep=T)
        s.generate(args.seed, n, 0)
        rule = native.RULE_BLAND if args.rule == "bland" else native.RULE_DANTZIG
        piv = 8
        o = native.make_opts(rule=rule, max_pivots=piv, loop_mode=native.LOOP_GRAPH, check_every=piv)
        s.run(o)                                    # warm-up (graph capture)
        s.generate(args.seed, n, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = s.run(o, hist_cap=piv)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        a5 = argparse.Namespace(rule=args.rule, seed=args.seed, rows=R, cols_total=C)
        hk = HistoryCheck(a5)
        hk.add(r)
        out = {"rows": R, "cols": C, "pivots": r["n_pivots"], "pivots_per_s": r["n_pivots"] / sec,
               "GBps": r["n_pivots"] / sec * 16.0 * R * C / 1e9, "loop": "rank-1 graph loop (k_update_tma)",
               "history": hk.report()}
        peak, _ = measured_peak()
        out["frac_of_measured_peak"] = out["GBps"] / peak
        s.generate(args.seed, n, 0)
        ob = native.make_opts(rule=rule, max_pivots=64, loop_mode=native.LOOP_BLOCKED, check_every=32)
        s.run(ob)
        s.generate(args.seed, n, 0)
        torch.cuda.synchronize()
        e0.record()
        rb = s.run(ob, hist_cap=64)
        e1.record()
        torch.cuda.synchronize()
        hb = HistoryCheck(a5)
        hb.add(rb)
        out["lookahead_K32"] = {"pivots_per_s": rb["n_pivots"] / (e0.elapsed_time(e1) * 1e-3), "history": hb.report()}
        s.close()
        del T
        torch.cuda.empty_cache()
        return out
    except Exception as exc:  # noqa: BLE001 -- an anchor, not the headline: report instead of failing the line
        torch.cuda.empty_cache()
        return {"skipped": f"{type(exc).__name__}: {exc}"}


def _solve_dense_host_ptr(solver, A_h, b_h, c_h, ops_h, opts, m, n, lda):
    """solve_dense on pinned torch host tensors without an extra numpy copy."""
    import ctypes as C
    from simplex_solver_b200 import native
    p = native.Problem(m, n, lda, C.c_void_p(A_h.data_ptr()), C.c_void_p(b_h.data_ptr()), C.c_void_p(c_h.data_ptr()),
                       ops_h.ctypes.data_as(C.c_void_p), 0, 0)
    res, keep = native.Solver._result(n, 0)
    native.check(native.lib().b200lp_solve_dense(solver._h, C.byref(p), C.byref(opts), C.byref(res)))
    return native.Solver._result_dict(res, keep, n)


def secondary_configs(args):
    """BASELINE configs 2 and 3 on this GPU next to the reference CPU path (scipy HiGHS, as solver_controller.py:78-85)."""
    import torch
    from simplex_solver_b200 import native, workloads as W
    out = {}
    s = native.Solver(0)
    # config 2: dense 1024 x 1024, Dantzig, end to end from host arrays
    A, b, c, ops, mx = W.dense_feasible_lp(1024, 0)
    cmin = -c
    s.solve_dense(A, b, cmin, ops)
    t0 = time.perf_counter()
    r = s.solve_dense(A, b, cmin, ops)
    gpu_s = time.perf_counter() - t0
    c2 = {"loop": "on-chip persistent kernel (tableau resident in shared memory, 1 grid barrier per pivot)",
          "gpu_e2e_s": gpu_s, "gpu_device_ms": r["device_ms"], "pivots": r["n_pivots"], "z": -r["fun"],
          "kernel_launches": r["kernel_launches"],
          "pivots_per_s": r["n_pivots"] / (r["device_ms"] * 1e-3), "us_per_pivot": r["device_ms"] * 1e3 / r["n_pivots"]}
    try:
        from scipy.optimize import linprog
        t0 = time.perf_counter()
        ref = linprog(cmin, A_ub=A, b_ub=b, bounds=[(0, None)] * 1024, method="highs-ds",
                      options={"presolve": True, "time_limit": 10})
        c2.update(reference_highs_s=time.perf_counter() - t0, reference_z=-ref.fun, reference_nit=int(ref.nit),
                  z_rel_err=abs(-r["fun"] + ref.fun) / abs(ref.fun), speedup_e2e=(time.perf_counter() - t0) / gpu_s,
                  reference_cores=1)
    except Exception as e:
        c2["reference_highs_s"] = f"unavailable: {e}"
    out["config2_dense1024_dantzig"] = c2
    # config 3: 100k x (20 x 30) batched (one GPU's view: the whole batch)
    import ctypes as C
    B = args.batch
    Ab, bb, cb, ob = W.batched_small_lps(0, B)
    m3, n3 = Ab.shape[1], Ab.shape[2]
    o = native.make_opts()
    # (a) kernel alone: inputs and outputs resident in HBM
    dev = [torch.from_numpy(a).cuda() for a in (Ab, bb, cb, ob)]
    dout = [torch.empty(B, dtype=torch.int32, device="cuda"), torch.empty(B, dtype=torch.float64, device="cuda"),
            torch.empty((B, n3), dtype=torch.float64, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda")]
    torch.cuda.synchronize()
    kernel_ms = min(s.solve_batched_device(B, m3, n3, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(),
                                           dev[3].data_ptr(), dout[0].data_ptr(), dout[1].data_ptr(), dout[2].data_ptr(),
                                           dout[3].data_ptr(), o) for _ in range(3))
    # (b) end to end from PINNED host buffers through the C ABI (H2D of A, b, c, ops and D2H of status, z, x, counts
    #     inside the call, chunked over two streams so that copies overlap the kernels)
    pin = [torch.from_numpy(a).pin_memory() for a in (Ab, bb, cb, ob)]
    outs = [torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.float64).pin_memory(),
            torch.empty((B, n3), dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()]
    ms = C.c_double()

    def call():
        native.check(native.lib().b200lp_solve_batched(
            s._h, B, m3, n3, C.c_void_p(pin[0].data_ptr()), C.c_void_p(pin[1].data_ptr()),
            C.c_void_p(pin[2].data_ptr()), C.c_void_p(pin[3].data_ptr()), C.byref(o), C.c_void_p(outs[0].data_ptr()),
            C.c_void_p(outs[1].data_ptr()), C.c_void_p(outs[2].data_ptr()), C.c_void_p(outs[3].data_ptr()), None, 0, 0,
            C.byref(ms)))
    call()
    walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        call()
        walls.append(time.perf_counter() - t0)
    wall_pinned = min(walls)
    rb = {"status": outs[0].numpy(), "fun": outs[1].numpy(), "n_pivots": outs[3].numpy()}
    assert (dout[0].cpu().numpy() == rb["status"]).all()
    h2d = Ab.nbytes + bb.nbytes + cb.nbytes + ob.nbytes
    c3 = {"batch": B, "kernel_ms": kernel_ms, "LPs_per_s_kernel": B / (kernel_ms * 1e-3),
          "LPs_per_s_e2e_pinned_host": B / wall_pinned, "e2e_pinned_ms": wall_pinned * 1e3, "h2d_bytes": int(h2d),
          "d2h_bytes": int(B * (4 + 8 + 8 * n3 + 4)), "pivots": int(rb["n_pivots"].sum()),
          "pivots_per_s_kernel": float(rb["n_pivots"].sum()) / (kernel_ms * 1e-3),
          "status_counts": {int(k): int(v) for k, v in zip(*np.unique(rb["status"], return_counts=True))}}
    try:
        from scipy.optimize import linprog
        k = 200
        t0 = time.perf_counter()
        agree = 0
        for i in range(k):
            A_ub = np.vstack([Ab[i][ob[i] == 0], -Ab[i][ob[i] == 1]])
            b_ub = np.concatenate([bb[i][ob[i] == 0], -bb[i][ob[i] == 1]])
            A_eq = Ab[i][ob[i] == 2]
            b_eq = bb[i][ob[i] == 2]
            ref = linprog(cb[i], A_ub=A_ub if len(b_ub) else None, b_ub=b_ub if len(b_ub) else None,
                          A_eq=A_eq if len(b_eq) else None, b_eq=b_eq if len(b_eq) else None,
                          bounds=[(0, None)] * Ab.shape[2], method="highs-ds", options={"presolve": True, "time_limit": 10})
            same = int(ref.status) == int(rb["status"][i])
            if same and ref.status == 0:
                same = abs(ref.fun - rb["fun"][i]) <= 1e-9 * max(1.0, abs(ref.fun))
            agree += bool(same)
        dt = time.perf_counter() - t0
        c3.update(reference_highs_LPs_per_s_1core=k / dt, reference_sample=k, reference_agree=agree)
    except Exception as e:
        c3["reference_highs_LPs_per_s_1core"] = f"unavailable: {e}"
    out["config3_batched_20x30"] = c3
    s.close()
    torch.cuda.empty_cache()
    return out


def opts_la(native, rule, n):
    return native.make_opts(rule=rule, max_pivots=n)


def bench_sharded(args):
    import torch
    import torch.distributed as dist
    from simplex_solver_b200 import native
    from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    with torch.cuda.stream(torch.cuda.Stream()):
        _bench_sharded(args, rank, local, world)
    dist.destroy_process_group()


def _bench_sharded(args, rank, local, world):
    import torch
    import torch.distributed as dist
    from simplex_solver_b200 import native
    from simplex_solver_b200.sharded import CudaShardEngine, ShardedTableau
    R = args.rows
    m = R - 1
    # ONE LP whatever the number of GPUs: n_total = cols_total - 1 structural variables, the same pivot sequence at
    # 1, 2, 4 and 8 GPUs.  Every shard stores its slice plus an RHS replica: cols_total / world columns, the last one
    # world - 1 more.
    n_total = args.cols_total - 1
    lo, hi = ShardedTableau.columns_of_tableau(args.cols_total, world, rank)
    ncols = hi - lo
    c_loc = ncols + 1                           # stored columns of this shard, RHS replica included
    rule = native.RULE_BLAND if args.rule == "bland" else native.RULE_DANTZIG
    opts = native.make_opts(rule=rule, max_pivots=args.pivots)

    # parity before anything is timed: the sharded loops on small tableaux against the oracle, on these GPUs
    parity = None
    if args.parity:
        sys.path.insert(0, ROOT)
        from tests.sharded_parity import run_cases
        n_cases, ok = run_cases(rank, world, local)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"cases": n_cases, "ok": bool(int(flag.item())), "ranks": world,
                  "checker": "oracle/ (pivot history, labels and every shard's tableau bits; tests/sharded_parity.py)"}

    eng = CudaShardEngine(m, n_total, lo, ncols, args.seed, device=local)
    exchange = args.exchange
    if exchange == "p2p":  # candidates stored straight into every peer's region over NVLink (no collective)
        # symmetric memory needs peer access between all GPUs of the node; if any rank cannot set it up, every rank
        # uses the NCCL all-gather exchange instead (same kernels either side of it) and the line says so
        try:
            eng.enable_p2p(world, rank)
            ok = 1
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] rank {rank}: peer-memory exchange unavailable ({exc}); using NCCL", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            eng.p2p = False
            exchange = "nccl"
    drv = ShardedTableau(eng, world, rank)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(args.warmup):
        drv.run(opts, args.pivots, check_every=args.pivots)
    # the timed steps start from the freshly generated tableau: their pivot history is a prefix of the sequence the
    # oracle computed for this config (tests/golden/pivot_history_config5.npy), identical at every number of GPUs
    eng.regenerate()
    torch.cuda.synchronize()
    hist = HistoryCheck(args)
    # rank 0 reads the NVLink counters through nvidia-smi (tens to hundreds of ms) BEFORE the barrier that starts the
    # timed region: read after it, the other ranks would start their clocks and then wait for rank 0 in the first exchange
    nvl0 = nvlink_counters(local) if rank == 0 else None
    dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pivots = 0
    for _ in range(args.steps):
        _, done_now = drv.run(opts, args.pivots, check_every=args.pivots)
        pivots += done_now
        if not os.environ.get("B200LP_BENCH_NO_HISTORY"):
            hist.add(eng.history(args.pivots))   # 3 x 4 x pivots bytes to the host, after the step's own status read
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    nvl1 = nvlink_counters(local) if rank == 0 else None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    value = pivots / sec
    bytes_per_pivot = 16.0 * R * args.cols_total
    if parity is not None:
        rep = hist.report()
        # every rank logged the same sequence?  (sum of per-rank "equals rank 0's hash" flags)
        mine = torch.tensor([int(rep["sha256"][:15], 16)], dtype=torch.int64, device=f"cuda:{local}")
        lo_h, hi_h = mine.clone(), mine.clone()
        dist.all_reduce(lo_h, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_h, op=dist.ReduceOp.MAX)
        rep["identical_on_all_ranks"] = bool(int(lo_h.item()) == int(hi_h.item()))
        parity["timed_history"] = rep
        parity["ok"] = parity["ok"] and rep["identical_on_all_ranks"] and rep.get("equals_oracle", True)

    # roofline of the update kernel on this shard, timed alone on the launching stream
    upd_ms = eng.solver.time_update(1, 1, native.UPDATE_AUTO, 5)
    shard_bytes = 16.0 * R * c_loc
    peak, peak_src = measured_peak()
    tt = torch.tensor([upd_ms], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    upd_ms = float(tt.item())
    achieved = shard_bytes / (upd_ms * 1e-3) / 1e9
    kname = "k_update_tma" if c_loc >= 32768 else "k_update_ldg<256,8>"   # b200lp.cu: resolve_variant (AUTO)
    # DRAM bytes per launch of the update kernel on this shard shape, from an ncu capture of one shard driven alone
    # (profiles/update_kernel_traffic.json -> "shards"; ncu cannot run under a multi-rank command)
    shard_traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "update_kernel_traffic.json")) as f:
            ent = json.load(f).get("shards", {}).get(f"{R}x{args.cols_total // world}")
        if ent:
            shard_traffic = ent["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": kname + " (per shard)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": shard_traffic,
                "algorithmic_bytes_per_launch": shard_bytes, "kernel_ms": {"update": upd_ms},
                "loop_GBps_aggregate": value * bytes_per_pivot / 1e9,
                "loop_frac_of_peak_per_gpu": value * bytes_per_pivot / 1e9 / world / peak,
                "loop_frac_of_nominal_8TBs_per_gpu": value * bytes_per_pivot / 1e9 / world / 8000.0}

    # beyond the rank-1 roofline: the look-ahead loop on the sharded tableau (same exchange per pivot, one tableau
    # pass per K pivots; bit-identical pivots, tests/nccl_sharded_check.py)
    lookahead = None
    if args.lookahead:
        lookahead = {}
        for K in (16, 32):
            n_la = 2 * K
            drv.run(opts_la(native, rule, n_la), n_la, check_every=K, lookahead=K)   # warm-up + graph capture
            eng.regenerate()
            torch.cuda.synchronize()
            dist.barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            _, got = drv.run(opts_la(native, rule, n_la), n_la, check_every=K, lookahead=K)
            a1.record()
            torch.cuda.synchronize()
            tl = torch.tensor([a0.elapsed_time(a1) * 1e-3], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
            pps = got / float(tl.item())
            hk = HistoryCheck(args)
            hk.add(eng.history(n_la))
            lookahead[f"K={K}"] = {"pivots_per_s": pps, "speedup_vs_rank1_loop": pps / value,
                                   "hbm_bytes_per_pivot": bytes_per_pivot / K,
                                   "roofline": lookahead_roofline(pps, R, args.cols_total, K, world), "history": hk.report()}
            if parity is not None:
                parity["ok"] = parity["ok"] and lookahead[f"K={K}"]["history"].get("equals_oracle", True)

    # e2e: every rank uploads its shard of the tableau from PINNED HOST memory inside the timed region (torch copy on
    # the solver's stream into the attached device tableau), runs the step and reads x*, z back.  The host copy of the
    # shard is made once, untimed (the generator is a device kernel; 137 GB cannot come from Python lists).  The step
    # is ONE solve call with the pivot budget of the one-GPU line's solve call (`e2e_pivots` = 512; the device-resident
    # steps above are 16 pivots only to last as long as the one-GPU step), so the upload weighs in the figure as it does
    # at N = 1 (a real solve of this LP takes > 10^5 pivots).  Falls back to regenerating the shard on the device -- and
    # says so -- when the host has no room to pin the tableau.
    e2e_pivots = max(args.pivots, args.e2e_pivots)
    shard_bytes_stored = 8 * eng.R * eng.ld
    host_ok = False
    avail = 0
    if args.e2e_host != "off":
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:  # noqa: BLE001
            avail = 0
        # all ranks pin their shards in the same host: the whole tableau (137 GB for config 5) plus working room
        host_ok = args.e2e_host == "on" or avail > 1.4 * 8 * R * args.cols_total
    hflag = torch.tensor([1 if host_ok else 0], dtype=torch.int32, device=f"cuda:{local}")
    dist.all_reduce(hflag, op=dist.ReduceOp.MIN)
    host_ok = bool(int(hflag.item()))
    host_T = None
    if host_ok:
        try:
            eng.regenerate()
            host_T = torch.empty(eng.R * eng.ld, dtype=torch.float64, pin_memory=True)
            host_T.copy_(eng.T)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] rank {rank}: pinned host copy of the shard failed ({exc}); e2e regenerates on the device", file=sys.stderr)
            host_T = None
        hflag = torch.tensor([1 if host_T is not None else 0], dtype=torch.int32, device=f"cuda:{local}")
        dist.all_reduce(hflag, op=dist.ReduceOp.MIN)
        if int(hflag.item()) == 0:
            host_T = None
    o_e2e = native.make_opts(rule=rule, max_pivots=e2e_pivots)
    rl0, cl0 = eng.initial_labels()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if host_T is not None:
        eng.T.copy_(host_T, non_blocking=True)      # H2D of the whole shard on the solver's stream
        eng.solver.set_labels(rl0, cl0)             # + its labels (host arrays)
    else:
        eng.regenerate()
    _, nq = drv.run(o_e2e, e2e_pivots, check_every=args.pivots)
    x, fun = eng.solution()
    e1.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e = {"value": nq / float(t2.item()) * bytes_per_pivot / 1e9, "unit": "GB/s", "pivots_per_s": nq / float(t2.item()),
           "h2d_bytes_per_step": int(shard_bytes_stored * world + 4 * world * (eng.R + eng.C)) if host_T is not None else 0,
           "d2h_bytes_per_step": int(8 * (n_total + 1) * world),
           "pivots_per_step": int(nq), "ms_per_step": float(t2.item()) * 1e3, "host_available_bytes": int(avail),
           "call": ("b200lp_attach'ed shard <- pinned host copy (H2D inside the timed region), sharded loop, "
                    "b200lp_read_solution -> host" if host_T is not None else
                    "shard regenerated on the device inside the timed region (host RAM cannot pin the 137 GB tableau), "
                    "sharded loop, b200lp_read_solution -> host")}
    if host_T is not None:
        hk = HistoryCheck(args)
        hk.add(eng.history(e2e_pivots))
        e2e["history"] = hk.report()
        if parity is not None:
            parity["ok"] = parity["ok"] and e2e["history"].get("equals_oracle", True)
    del host_T
    # BASELINE config 3 at N GPUs: the 100k independent 20 x 30 LPs in contiguous blocks per rank (generator blocks of
    # 1000), one warp per LP, no data-path collective; timed on the device, max over ranks
    batched = None
    if args.secondary:
        batched = bench_batched_sharded(args, rank, local, world, eng.solver)
    # per pivot and rank: the fused pick kernel + the update kernel (peer-memory exchange); the all-gather exchange has
    # price, extract, [NCCL], winner, ratio, update
    launches = args.steps * args.pivots * (2 if exchange == "p2p" else 5)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value * bytes_per_pivot / 1e9, "unit": "GB/s", "pivots_per_s": value,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "roofline": roofline,
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "collective": {"op": ("peer-memory exchange inside the pick kernel: k_shard_pick stores every candidate into every "
                                  "peer's region over NVLink and polls its own (torch symmetric memory; no collective call)"
                                  if exchange == "p2p" else "all_gather_into_tensor (NCCL)"),
                           "bytes_per_rank_per_pivot": 8 * (R + 2),
                           "algorithmic_nvlink_bytes_per_rank_per_pivot": 8 * (R + 2) * (world - 1),
                           "nvlink_counters_rank0": None if not (nvl0 and nvl1) else {
                               "tx_bytes_per_pivot": (nvl1[0] - nvl0[0]) / max(pivots, 1),
                               "rx_bytes_per_pivot": (nvl1[1] - nvl0[1]) / max(pivots, 1),
                               "source": nvl1[2] + " of GPU 0, difference over the timed region"}},
        }
        if parity is not None:
            line["parity"] = parity
        if lookahead:
            line["lookahead"] = lookahead
        if batched:
            line["secondary"] = {"config3_batched_20x30": batched}
        print(json.dumps(line))
    if parity is not None and not parity["ok"]:
        if rank == 0:
            print("[bench] PARITY FAILURE: the sharded GPU path differs from the oracle", file=sys.stderr)
        dist.destroy_process_group()
        sys.exit(3)


def bench_batched_sharded(args, rank, local, world, solver):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from simplex_solver_b200 import native
    from simplex_solver_b200 import workloads as W
    from simplex_solver_b200.batched import shard_range
    B = args.batch
    lo, hi = shard_range(B, world, rank, W.BATCH_BLOCK)
    k = hi - lo
    A, b, c, ops = W.batched_small_lps(lo, k)
    m3, n3 = A.shape[1], A.shape[2]
    o = native.make_opts()
    dev = [torch.from_numpy(a).to(f"cuda:{local}") for a in (A, b, c, ops)]
    dout = [torch.empty(k, dtype=torch.int32, device=f"cuda:{local}"), torch.empty(k, dtype=torch.float64, device=f"cuda:{local}"),
            torch.empty((k, n3), dtype=torch.float64, device=f"cuda:{local}"), torch.empty(k, dtype=torch.int32, device=f"cuda:{local}")]

    def kernel():
        return solver.solve_batched_device(k, m3, n3, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(),
                                           dev[3].data_ptr(), dout[0].data_ptr(), dout[1].data_ptr(), dout[2].data_ptr(),
                                           dout[3].data_ptr(), o)
    kernel()
    torch.cuda.synchronize()
    dist.barrier()
    ms = min(kernel() for _ in range(3))
    pin = [torch.from_numpy(a).pin_memory() for a in (A, b, c, ops)]
    outs = [torch.empty(k, dtype=torch.int32).pin_memory(), torch.empty(k, dtype=torch.float64).pin_memory(),
            torch.empty((k, n3), dtype=torch.float64).pin_memory(), torch.empty(k, dtype=torch.int32).pin_memory()]
    dms = C.c_double()

    def call():
        native.check(native.lib().b200lp_solve_batched(
            solver._h, k, m3, n3, C.c_void_p(pin[0].data_ptr()), C.c_void_p(pin[1].data_ptr()),
            C.c_void_p(pin[2].data_ptr()), C.c_void_p(pin[3].data_ptr()), C.byref(o), C.c_void_p(outs[0].data_ptr()),
            C.c_void_p(outs[1].data_ptr()), C.c_void_p(outs[2].data_ptr()), C.c_void_p(outs[3].data_ptr()), None, 0, 0,
            C.byref(dms)))
    call()
    walls = []
    for _ in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        call()
        walls.append(time.perf_counter() - t0)
    same = bool((dout[0].cpu().numpy() == outs[0].numpy()).all())
    t = torch.tensor([ms, min(walls) * 1e3, 0.0 if same else 1.0, float(outs[3].numpy().sum())], dtype=torch.float64,
                     device=f"cuda:{local}")
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    kernel_ms, e2e_ms, bad, pivots = float(tmax[0]), float(tmax[1]), float(tmax[2]), float(t[3])
    return {"batch": B, "per_rank": k, "kernel_ms_max_over_ranks": kernel_ms, "LPs_per_s_kernel": B / (kernel_ms * 1e-3),
            "pivots_per_s_kernel": pivots / (kernel_ms * 1e-3), "e2e_pinned_ms_max_over_ranks": e2e_ms,
            "LPs_per_s_e2e_pinned_host": B / (e2e_ms * 1e-3), "device_vs_host_path_status_equal": bad == 0.0,
            "partition": "contiguous blocks of the batch per rank, no collective in the solve (SURVEY 8e)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pivots", type=int, default=None, help="pivots per step")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--cols-total", type=int, default=None)
    ap.add_argument("--rule", default="bland", choices=["bland", "dantzig"])
    ap.add_argument("--variant", default="auto", choices=["auto", "ldg", "tma"])
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--batch", type=int, default=100000)
    ap.add_argument("--ref-pivots", type=int, default=None,
                    help="pivots per step of the CPU reference arm (default 32 at N=1, 8 on the config-5 slab)")
    ap.add_argument("--cpu-pivots", type=int, default=384, help="pivots of the cpu_baseline sample")
    ap.add_argument("--e2e-pivots", type=int, default=512,
                    help="N > 1: pivot budget of the end-to-end solve call (default: the one-GPU line's 512)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
    ap.add_argument("--no-lookahead", dest="lookahead", action="store_false")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="per-pivot exchange of the sharded loops")
    ap.add_argument("--generator", default="counter", choices=["counter", "survey"],
                    help="N = 1 workload: the counter-based device generator shared with the sharded config (default) or "
                         "SURVEY 8d's literal config-2 generator with seed 4, built on the host and uploaded")
    ap.add_argument("--no-parity", dest="parity", action="store_false",
                    help="skip the parity block (small cases against the oracle before the timed region)")
    ap.add_argument("--no-config5-one-gpu", dest="config5_one_gpu", action="store_false",
                    help="N = 1: skip the 137 GB config-5 tableau on one GPU (same-workload anchor of the 1 -> 8 curve)")
    ap.add_argument("--e2e-host", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: end-to-end step from pinned HOST copies of the shards (auto: when host RAM allows)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        args.gpus = world
    if args.gpus == 1:
        args.rows = args.rows or 16384
        args.cols_total = args.rows
        args.pivots = args.pivots or 512
        args.ref_pivots = args.ref_pivots or 32
    else:
        args.rows = args.rows or 131072
        args.cols_total = args.cols_total or 131072
        args.pivots = args.pivots or 16
        args.ref_pivots = args.ref_pivots or 8
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.gpus == 1:
        bench_single_gpu(args)
    else:
        bench_sharded(args)


if __name__ == "__main__":
    main()
