#!/usr/bin/env python
"""bench.py -- restarts/sec of the batched factorize hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic 20 000 cells x 2 000 HVG fp32, K=10, 100 restarts,
solver 'mu' (Frobenius), tol 1e-4, max_iter 1000 -- the reference's own restart loop
(cnmf.py:735-745) with its own seed rule (cnmf.py:597-610).  One "step" = factorizing the whole
batch of 100 restarts to convergence.  With N > 1 ranks (torchrun) every rank holds a replica of X
and factorizes its OWN 100 restarts (weak scaling: 100*N restarts in total, job split idx % N == rank as
in cnmf.py:52-53) and the step ends with the NCCL all-gather of spectra that replaces `combine`.

Printed JSON (rank 0): see the task contract; `value` = restarts/s with X and the initial factors
resident in HBM, `e2e` = the same through the plugin call with HOST buffers (H2D of X and of the
host-RNG initial factors, D2H of the spectra inside the timed region).
`--impl reference` times the reference's CPU implementation (oracle/reference_path.py: the
reference's call sequence on scikit-learn, float64) on a bounded sample -- one restart per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NMF restarts/sec on NxG counts, K-sweep x100 seeds, 1/2/4/8 B200 vs sklearn CPU"
N_CELLS, N_GENES, K, N_RESTARTS = 20000, 2000, 10, 100
NMF_KW = dict(solver="mu", beta_loss=2.0, tol=1e-4, max_iter=1000, init="random", alpha_W=0.0, alpha_H=0.0, l1_ratio=0.0)


def workload_config(world):
    return {"workload": "c2: synthetic %dx%d fp32 (Poisson counts / gene std), K=%d, %d restarts per GPU, solver=mu "
                        "(Frobenius), tol=1e-4, max_iter=1000" % (N_CELLS, N_GENES, K, N_RESTARTS),
            "restarts_total": N_RESTARTS * world, "parallelism": "restarts sharded x%d, X replicated" % world,
            "l2": "inputs larger than L2 (X forms 640 MB, factors 264 MB)"}


def make_data():
    from cnmf_b200.synth import make_counts, normalise
    counts = make_counts(N_CELLS, N_GENES, k_true=12, seed=0)
    X, _ = normalise(counts, np.float32)
    return X


def restart_jobs(world, rank):
    from cnmf_b200.synth import restart_table
    rows = restart_table([K], N_RESTARTS * world, seed=14)
    mine = [rows[i] for i in range(len(rows)) if (i - rank) % world == 0]
    return [r[0] for r in mine], [r[2] for r in mine]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [c for c in sm if c > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        m = json.load(open(p))
        return m.get("bf16_tflops_sustained", m.get("bf16_tflops")) / 2.0, \
            "MEASURED_PEAKS.json bf16_tflops_sustained/2 (dense TF32 = half the bf16 rate), of measured"
    return 1400.0 / 2.0, "fallback 1.4 PFLOP/s sustained bf16 / 2, of fallback"


def measured_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth), of measured"
    return 6500.0, "fallback 6.5 TB/s copy bandwidth, of fallback"


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import reference_path
    try:      # torchrun exports OMP_NUM_THREADS=1 to its children: give the reference all the host threads back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    X = make_data().astype(np.float64)
    ks, seeds = restart_jobs(1, 0)
    jobs = list(zip(ks, seeds))
    n = args.steps + args.warmup
    for i in range(args.warmup):
        reference_path.factorize(X, [jobs[i % len(jobs)]], "mu")
    t0 = time.perf_counter()
    its = []
    for i in range(args.steps):
        _, it, _ = reference_path.factorize(X, [jobs[(args.warmup + i) % len(jobs)]], "mu")
        its += it
    dt = time.perf_counter() - t0
    val = args.steps / dt
    try:
        from threadpoolctl import threadpool_info
        threads = max([t.get("num_threads", 1) for t in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count()
    sample = "1 restart per step (K=%d, reference seeds #%d..), sklearn non_negative_factorization MU float64 to convergence, n_iter=%s" % (
        K, args.warmup, its)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "restarts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(1),
        "cpu_baseline": {"value": val, "unit": "restarts/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "restarts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_ours(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from cnmf_b200 import _lib
    from cnmf_b200.engine import Engine
    from cnmf_b200.parallel import init_process_group
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cnmf_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    init_process_group("nccl")
    dev = torch.device("cuda:%d" % local)

    X = make_data()
    Xpin = torch.from_numpy(X).pin_memory()
    Xnp = Xpin.numpy()
    ks, seeds = restart_jobs(world, rank)
    SK = int(sum(ks))
    eng = Engine(local)
    lib = _lib.load()

    # ---------------- device-resident arm: X and the initial factors already in HBM ----------------
    ds = eng.dataset(Xnp, precision=args.precision)
    passes = 2 if ds.exact else 3
    f16 = bool(ds.f16)
    ld_r, ld_c = ds.ld()
    s, _ = ds.sums()
    mean = s / (N_CELLS * float(X.shape[1]))
    W0 = np.zeros((SK, ld_r), np.float32)
    H0 = np.zeros((SK, ld_c), np.float32)
    o = 0
    for k, seed in zip(ks, seeds):
        _lib.check(lib.cnmf_random_init_host(seed, float(np.sqrt(mean / k)), X.shape[0], X.shape[1], k,
                                             _lib.ptr(W0[o:o + k]), ld_r, _lib.ptr(H0[o:o + k]), ld_c))
        o += k
    W0_t = torch.from_numpy(W0).to(dev)
    H0_t = torch.from_numpy(H0).to(dev)
    out_t = torch.empty((SK, ld_c), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * SK, ld_c), dtype=torch.float32, device=dev) if world > 1 else None

    def step_resident():
        n_iter, _ = ds.factorize_dev(ks, W0_t.data_ptr(), H0_t.data_ptr(), out_t.data_ptr(), NMF_KW)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_t)       # the single collective: spectra all-gather
        return n_iter

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local)
    sync_all()
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    n_iter = None
    for _ in range(args.steps):
        n_iter = step_resident()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    gemm_ms, gemm_launches, gemm_flops = eng.profile_get(0)
    upd_ms, upd_launches, upd_bytes = eng.profile_get(1)
    eng.profile(False)
    launches = eng.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = N_RESTARTS * world * args.steps / (ms * 1e-3)

    # ---------------- end-to-end arm: host buffers through the plugin call ----------------
    ds.close()

    phases = {"dataset_ms": 0.0, "rng_ms": 0.0, "h2d_ms": 0.0, "solve_ms": 0.0, "d2h_ms": 0.0}

    def step_e2e():
        t_ds = time.perf_counter()
        d2 = eng.dataset(Xnp, precision=args.precision)         # H2D of X + device-side prep
        phases["dataset_ms"] += 1e3 * (time.perf_counter() - t_ds)
        sp, _, it, _ = d2.factorize(ks, seeds, NMF_KW)           # host RNG init, H2D, solve, D2H of spectra
        for k_, v_ in eng.last_timing().items():
            phases[k_] += v_
        d2.close()
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_t)
        return sp

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    sync_all()
    for k_ in phases:
        phases[k_] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    sync_all()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N_RESTARTS * world * args.steps / float(t.item())
    clk = clocks.stop() if rank == 0 else None      # sampled across both timed regions (resident + end-to-end)
    h2d = X.shape[0] * X.shape[1] * 4 + SK * (ld_r + ld_c) * 4
    d2h = SK * X.shape[1] * 4

    if rank != 0:
        return

    # ---------------- informational: the consensus stage (cnmf.py:871-919) on this step's spectra ----------------
    consensus = None
    try:
        from cnmf_b200 import consensus as cs
        d3 = eng.dataset(Xnp, precision=args.precision)
        sp, _, _, _ = d3.factorize(ks, seeds, NMF_KW)
        merged = np.vstack(sp)
        torch.cuda.synchronize(dev)
        tc = [time.perf_counter()]
        S = cs.SpectraMatrix(eng, merged).l2_normalize()
        dens, _ = S.local_density(int(0.3 * merged.shape[0] / K))
        tc.append(time.perf_counter())
        keep = np.where(dens < 0.5)[0]
        S2 = S.take_rows(keep) if len(keep) < S.R else S
        labels, labels_t, _, _ = cs.kmeans(S2, K)
        tc.append(time.perf_counter())
        med = cs.cluster_medians(S2, labels_t, K)
        tc.append(time.perf_counter())
        W, it_refit, _ = d3.refit(med, NMF_KW)
        tc.append(time.perf_counter())
        d3.close()
        consensus = {"R": int(merged.shape[0]), "kept": int(len(keep)), "ms": {
            "upload_l2_density": 1e3 * (tc[1] - tc[0]), "kmeans_n_init10": 1e3 * (tc[2] - tc[1]),
            "cluster_median": 1e3 * (tc[3] - tc[2]), "refit_usage": 1e3 * (tc[4] - tc[3])},
            "refit_n_iter": int(it_refit), "note": "wall clock, outside the timed factorize regions"}
    except Exception as ex:          # never let the informational block break the bench line
        consensus = {"error": repr(ex)}
    peak, peak_src = measured_peaks()
    if f16:           # kind::f16 runs at the bf16 rate: the denominator is the measured bf16 figure itself
        peak, peak_src = 2.0 * peak, peak_src.replace("/2 (dense TF32 = half the bf16 rate)", " (kind::f16 = the bf16 rate)")
    hbm_peak, hbm_src = measured_hbm()
    upd_gbs = upd_bytes / (upd_ms * 1e-3) / 1e9 if upd_ms > 0 else 0.0
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("dram_bytes_per_launch_f16" if f16 else "dram_bytes_per_launch", tj.get("dram_bytes_per_launch"))
    out = {
        "metric": METRIC, "value": value, "unit": "restarts/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": ("f32 (2-pass split-fp16 tensor-core products of row-normalised factors x exact integer counts, fp32 "
                  "accumulate)" if f16 else "f32 (%d-pass split-TF32 tensor-core products, fp32 accumulate)" % passes),
        "data": "synthetic",
        "config": workload_config(world),
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "restarts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": {k_: v_ / args.steps for k_, v_ in phases.items()}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tf32x3_kernel<256,%s>" % ("3,exact-B,kind::f16" if f16 else "3,exact-B" if passes == 2 else "2,general"),
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "mma_passes": passes, "mma_frac": passes * achieved / peak,
                     "note": "achieved = algorithmic 2*M*N*K per launch (counted once, not %dx for the TF32 passes) / "
                             "CUDA-event launch time; %d launches, %.1f ms of %.1f ms timed; peak = %s; X %s" % (
                                 passes, gemm_launches, gemm_ms, ms, peak_src,
                                 "recognised as scaled integer counts -> exact B operand, 2 passes" if passes == 2
                                 else "general real matrix -> 3 passes")},
        "roofline_update": {"bound": "hbm", "kernel": "update_kernel<16,mu> (multiplicative update fused with the Gram "
                            "of the factor it writes and the emission of its tensor-core operand pieces)", "achieved": upd_gbs, "peak": hbm_peak, "unit": "GB/s",
                            "frac": upd_gbs / hbm_peak, "traffic": None,
                            "note": "second kernel of the step: achieved = algorithmic bytes per launch (factor read + "
                                    "product slices read + factor and its 2 operand pieces written: 2 x fp16 or 2 x tf32, x live rows x items) / "
                                    "CUDA-event launch time; %d launches, %.1f ms of %.1f ms timed; peak = %s" % (
                                        upd_launches, upd_ms, ms, hbm_src)},
        "n_iter": {"mean": float(np.mean(n_iter)), "max": int(np.max(n_iter))},
        "consensus": consensus,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import reference_path
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=os.cpu_count())
        except Exception:
            pass
        _, its, sec = reference_path.factorize(X.astype(np.float64), [(ks[0], seeds[0])], "mu")
        try:
            from threadpoolctl import threadpool_info
            threads = max([t.get("num_threads", 1) for t in threadpool_info()] + [1])
        except Exception:
            threads = os.cpu_count()
        out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "restarts/s", "cores": threads, "kind": "port",
                               "sample": "1 of the 100 restarts (K=%d, seed %d) to convergence (n_iter=%d), sklearn "
                                         "non_negative_factorization MU float64 as called by cnmf.py:672" % (ks[0], seeds[0], its[0])}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", type=str, default="f16x2", choices=["f16x2", "tf32x3", "tf32x3-general", "fp32"],
                    help="f16x2 (default): 2 kind::f16 passes when X is scaled integer counts, else 3 kind::tf32 passes; tf32x3: 2 / 3 kind::tf32 passes")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local)
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
