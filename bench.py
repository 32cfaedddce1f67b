#!/usr/bin/env python
"""bench.py — orbital-optimisation energy+gradient evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3]): synthetic 8-fold-symmetric ERI, M=256 spatial orbitals, N=16
active orbitals, ensemble-N-representable RDMs, FP64.  One step = one (E, dE/dU) evaluation at a
fresh partial unitary U.  With N GPUs the ERI tensor is sharded by its first index (strong scaling
of the same problem) and each evaluation ends with one all-reduce of M*N+1 doubles.

Prints ONE JSON line (rank 0):
  value            evaluations/s with U already in HBM, K steps timed with CUDA events (a BURST of
                   K*ms_per_step; `sustained` is the same loop run for >= 2 s with the SM clock)
  e2e              the same through the host-buffer entry points (pinned H2D of U, D2H of E and
                   dE/dU every step, pipelined two deep: oo_eval_submit / oo_eval_wait)
  roofline         dominant kernel (K1: TMA + DMMA half-transform with the fused 2-RDM contraction)
                   against the measured HBM copy peak; frac_read_peak = against the read-only
                   stream measured live (K1 only reads)
  roofline_tensor  K1 against the DMMA peak: frac = vs the peak measured after the loop (power
                   capped), frac_cold = vs the peak measured on the idle GPU before it
  parity           E and dE/dU of step 0 (all-reduced) against the fixture the 1-GPU run produced
  config5          BASELINE.json configs[4] (M=400, N=24) in the same invocation: dense first-index
                   shard when N >= 2, pair-packed on one GPU
  cpu_baseline     the reference's torch formulation (oracle/torch_port.py) on this box's host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_BENCH, N_BENCH = 256, 16
METRIC = "orbital-opt energy+grad evals/sec at M=256,N=16 (FP64)"
WORKLOAD = "synthetic 8-fold-symmetric ERI M=256, N=16 spatial, random N-representable RDMs, FP64"
PARITY_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_parity_M256_N16.npz")


def _load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _load_traffic(M, N, mloc, slabs):
    """Per-launch DRAM bytes of K1 from the committed ncu capture, if it is for this shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            t = json.load(f)
        if (t["M"], t["N"], t["mloc"], t.get("slabs", mloc * M)) == (M, N, mloc, slabs):
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=25):
        self.index, self.proc, self.lines, self.period_ms = index, None, [], period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, window=None):
        """Median SM clock, throttle reasons and peak power of the samples taken inside `window`
        (host time stamps (t0, t1) of the timed region; nvidia-smi is started long before it so
        that it is already streaming), or of all samples."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        picked = [ln for (ts, ln) in self.lines
                  if window is None or window[0] <= ts <= window[1] + 0.5 * self.period_ms * 1e-3]
        if window is not None and not picked and self.lines:      # region shorter than a period
            mid = 0.5 * (window[0] + window[1])
            picked = [min(self.lines, key=lambda tl: abs(tl[0] - mid))[1]]
        for ln in picked:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


def _window_clocks(sampler, window):
    """Parse the samples of an already stopped sampler for another window."""
    class _Done:                      # stop() terminates the process first: give it a finished one
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

        def kill(self):
            pass
    sampler.proc = _Done()
    return sampler.stop(window)


def _host_threads():
    """All the host cores for the CPU baseline: torch.distributed.run exports OMP_NUM_THREADS=1,
    which would starve it (round-1 SCALE ratios at N >= 2 were 9x too high for that reason)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_baseline(target_seconds=12.0, allow_full=True):
    """The reference's torch formulation (einsum forward + autograd backward, left-to-right
    contraction order) on the host cores.  When host memory allows, ONE FULL M=256 evaluation on
    the whole 34 GB spatial tensor is timed; else (or in addition) a slab of the last ERI index,
    in which every contraction of the chain is linear."""
    import torch
    import esoo_b200  # noqa: F401  (synthetic generators live in the package)
    from esoo_b200 import synthetic
    from oracle import torch_port
    M, N = M_BENCH, N_BENCH
    cores = _host_threads()
    h = synthetic.h_spatial(M)
    D, G = synthetic.rdms_spatial(N)
    U = synthetic.random_partial_unitary(M, N)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    out = {"unit": "evals/s", "cores": cores, "kind": "port"}
    ms = 32
    g = synthetic.eri_spatial_shard(M, 0, ms).permute(3, 2, 1, 0).contiguous()
    torch_port.time_reference(U, D, G, h, g[..., :4].contiguous(), 0)          # warm-up
    t_eval, t_iter = torch_port.time_reference(U, D, G, h, g, 0, repeats=1)
    reps = max(1, min(20, int(target_seconds / max(t_iter, 1e-3))))
    t_eval, t_iter = torch_port.time_reference(U, D, G, h, g, 0, repeats=reps)
    frac = ms / M
    out.update({"value": frac / t_eval,
                "sample": (f"slab of {ms}/{M} of the last ERI index (all contractions of the chain "
                           f"are linear in it), {reps} repeats of einsum forward + autograd "
                           f"backward, {t_eval:.3f} s each; spatial M^4 tensor = 32x less work "
                           f"than the reference's spin-orbital tensor"),
                "reference_iterations_per_s": frac / t_iter})
    del g
    if allow_full and avail > 90 * (1 << 30):
        t_a = time.perf_counter()
        g_full = synthetic.eri_spatial(M)                  # 34.4 GB on the host
        t_b = time.perf_counter()
        # two repeats, the faster one counts (the first also pays the page faults of the
        # intermediates): the baseline gets its best
        runs = [torch_port.time_reference(U, D, G, h, g_full, 0, repeats=1) for _ in range(2)]
        t_eval_full, t_iter_full = min(r[0] for r in runs), min(r[1] for r in runs)
        del g_full
        out["slab_extrapolated_value"] = out["value"]
        out["value"] = 1.0 / t_eval_full
        out["reference_iterations_per_s"] = 1.0 / t_iter_full
        out["sample"] = (f"FULL evaluation on the whole M={M} spatial tensor (34.4 GB, built "
                         f"in {t_b - t_a:.1f} s), best of 2: einsum forward + autograd backward "
                         f"{t_eval_full:.2f} s; the {ms}/{M} slab sample extrapolates to "
                         f"{out['slab_extrapolated_value']:.3f} evals/s; spatial tensor = 32x less "
                         f"work than the reference's spin-orbital tensor")
    else:
        out["sample"] += f"; full-tensor run skipped (host memory available {avail / 2**30:.0f} GiB)"
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    # one bounded sample per run: the slab of the last ERI index (repeated for ~8 s) and, when the
    # host has the memory, one full evaluation on the whole tensor, which then is the value
    base = cpu_baseline(target_seconds=8.0, allow_full=True)
    value = base["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "M": M_BENCH, "N": N_BENCH},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def build_engine(M, N, dev, rank, world, args, packed, dense=False):
    """Engine holding this rank's shard of the synthetic ERI tensor (+ communicators)."""
    import torch
    import torch.distributed as dist
    import esoo_b200
    from esoo_b200 import synthetic
    t0, mloc = esoo_b200.shard_range(M, rank, world)
    h = synthetic.h_spatial(M, device=dev)
    D, G = synthetic.rdms_spatial(N)
    eng = esoo_b200.OrbitalEngine(M, N, device=dev, t0=t0, mloc=mloc)
    if packed:
        g = synthetic.eri_spatial_pair_packed(M, t0, mloc, device=dev)
        eng.set_integrals_packed(h, g)         # symmetric by construction (half is not stored)
    else:
        g = synthetic.eri_spatial_shard(M, t0, mloc, device=dev)
        if world == 1:
            eng.set_integrals(h, g)            # verifies the V4 symmetry on the device
        else:
            eng.set_integrals(h, g, assume_v4_symmetric=True)   # symmetric by construction
    mode = "none (1 GPU)"
    if world > 1:
        esoo_b200.attach_nccl(eng)
        mode = "NCCL all-reduce of M*N+1 doubles"
        if args.allreduce == "fused":
            try:
                esoo_b200.attach_peer_memory(eng)
                ok = 1
            except Exception as exc:            # no peer access on this box: NCCL still works
                if rank == 0:
                    print(f"# fused all-reduce unavailable ({exc}); using NCCL", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                mode = "one-shot all-reduce fused into k_tail_reduce over NVLink peer memory"
            elif ok:
                raise SystemExit("ranks disagree on the all-reduce mode")
    eng.set_rdms(D, G)
    if not packed:
        eng.set_pair_symmetry(not dense)
    return eng, g, t0, mloc, mode


def timed_evals(eng, U_dev, K, W, stream, barrier, world, dev):
    """K evaluations timed with CUDA events on the library's stream; max over ranks (ms)."""
    import torch
    import torch.distributed as dist
    for i in range(W):
        eng.enqueue_energy_grad(U_dev[i % len(U_dev)])
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for i in range(K):
            eng.enqueue_energy_grad(U_dev[(W + i) % len(U_dev)])
        ev1.record(stream)
    barrier()
    t_host1 = time.time()
    tt = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item()), (t_host0, t_host1)


def k1_timing(eng, U_dev, K, W, world, dev):
    import torch
    import torch.distributed as dist
    eng.set_timing(True)
    k1_ms, parts = [], [0.0] * 5
    for i in range(K):
        eng.enqueue_energy_grad(U_dev[(W + i) % len(U_dev)], allreduce=False)
        t = eng.last_timing()
        k1_ms.append(t[0])
        parts = [a + b for a, b in zip(parts, t)]
    eng.set_timing(False)
    kt = torch.tensor([sum(k1_ms) / len(k1_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    return float(kt.item()), k1_ms, [p / K for p in parts]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--M", type=int, default=M_BENCH)
    ap.add_argument("--N", type=int, default=N_BENCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the M=400, N=24 block")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained block")
    ap.add_argument("--write-parity-fixture", action="store_true",
                    help="(1 GPU) store E and dE/dU of step 0 as the fixture the other N compare to")
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU: all-reduce fused into the tail kernel over NVLink peer "
                         "memory (default) or a separate NCCL call")
    ap.add_argument("--dense", action="store_true",
                    help="stream every slab of the shard instead of one per (t,q)/(q,t) pair")
    ap.add_argument("--packed", action="store_true",
                    help="pair-packed ERI storage (OO_G_PAIR_PACKED): only the streamed slab of "
                         "every (t,q)/(q,t) pair is resident, half the memory (M=400 fits one GPU)")
    args = ap.parse_args()
    if args.packed and args.dense:
        raise SystemExit("--packed stores only the pair-symmetric slab set; it excludes --dense")
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import esoo_b200
    from esoo_b200 import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    M, N, K, W = args.M, args.N, args.steps, max(3, args.warmup)
    headline = (M, N) == (M_BENCH, N_BENCH)
    # nvidia-smi takes a few hundred ms to deliver its first sample: start it now, filter its samples
    # by the host time stamps of each timed region later
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # FP64 peaks of the idle GPU, before anything heats it up
    cold = esoo_b200.measure_peaks(local, 4 << 30) if rank == 0 else None
    eng, g, t0, mloc, ar_mode = build_engine(M, N, dev, rank, world, args, args.packed, args.dense)
    slabs = eng.streamed_slabs()
    stream = torch.cuda.Stream(device=dev)
    eng.use_stream(stream)

    # a fresh U per step (identical on every rank), resident in HBM for `value`,
    # in pinned host memory for `e2e`
    Us = [synthetic.random_partial_unitary(M, N, seed=synthetic.SEED_U + i) for i in range(K + W)]
    U_dev = [u.to(dev) for u in Us]
    U_host = [u.numpy() for u in Us]
    torch.cuda.synchronize()

    # ---------------- parity of step 0 (all-reduced E and dE/dU) -------------------------------
    E0, g0 = eng.energy_grad(U_dev[W])
    g0 = g0.cpu().numpy()
    parity = {"E_step0": float(E0), "grad_l2_step0": float(np.linalg.norm(g0)),
              "grad_abs_sum_step0": float(np.abs(g0).sum())}
    if headline and rank == 0:
        if args.write_parity_fixture and world == 1:
            np.savez_compressed(PARITY_FIXTURE, E=float(E0), grad=g0, M=M, N=N, step=W,
                                seed=synthetic.SEED_U + W)
        if os.path.isfile(PARITY_FIXTURE):
            ref = np.load(PARITY_FIXTURE)
            parity.update({
                "against": "tests/golden/bench_parity_M256_N16.npz (the 1-GPU run's step 0)",
                "dE_rel": abs(float(E0) - float(ref["E"])) / max(1.0, abs(float(ref["E"]))),
                "dgrad_rel": float(np.linalg.norm(g0 - ref["grad"]) / np.linalg.norm(ref["grad"])),
                "tolerance": {"dE_rel": 1e-10, "dgrad_rel": 1e-9}})
            parity["ok"] = parity["dE_rel"] <= 1e-10 and parity["dgrad_rel"] <= 1e-9

    # ---------------- device-resident throughput (burst) ---------------------------------------
    launches0 = eng.launch_count()
    ms_total, window = timed_evals(eng, U_dev, K, W, stream, barrier, world, dev)
    launches_per_eval = (eng.launch_count() - launches0) // (K + W)
    value = K / (ms_total * 1e-3)

    # ---------------- per-kernel timing of the dominant kernel (CUDA events, same stream) -----
    k1_avg, k1_ms, parts = k1_timing(eng, U_dev, K, W, world, dev)

    # ---------------- sustained: the same loop for >= 2 s ---------------------------------------
    sustained = None
    if not args.no_sustained:
        n_sus = max(K, int(2.2 / (ms_total * 1e-3 / K)))
        ms_sus, window_sus = timed_evals(eng, U_dev, n_sus, W, stream, barrier, world, dev)
        sustained = {"value": n_sus / (ms_sus * 1e-3), "unit": "evals/s", "steps": n_sus,
                     "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus,
                     "window": window_sus}

    # ---------------- end to end through host buffers (pipelined two deep) ---------------------
    for i in range(W):
        eng.energy_grad_host(U_host[i])
    barrier()
    t_a = time.perf_counter()
    e_sum = 0.0
    eng.submit_host(U_host[W], 0)
    for i in range(K):
        if i + 1 < K:
            eng.submit_host(U_host[W + i + 1], (i + 1) & 1)
        e, _grad = eng.wait_host(i & 1)
        e_sum += e
    torch.cuda.synchronize()
    t_b = time.perf_counter()
    e2e_t = torch.tensor([t_b - t_a], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = K / float(e2e_t.item())
    # the synchronous call (one evaluation in flight) for comparison
    barrier()
    t_a = time.perf_counter()
    for i in range(K):
        eng.energy_grad_host(U_host[W + i])
    t_b = time.perf_counter()
    sync_t = torch.tensor([t_b - t_a], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sync_t, op=dist.ReduceOp.MAX)
    e2e_sync_value = K / float(sync_t.item())

    # ---------------- whole inner loop on the device (eval + all-reduce + retraction + BB + stop) --
    n_inner = 60
    barrier()
    t_c = time.perf_counter()
    res = eng.optimize(U_host[0], 1e-3, 0.0, n_inner)      # tol=0: runs until iteration > maxiter
    t_d = time.perf_counter()
    inner_t = torch.tensor([t_d - t_c], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(inner_t, op=dist.ReduceOp.MAX)
    inner_iters = res["n_iter"]
    inner_rate = inner_iters / float(inner_t.item())
    hot = esoo_b200.measure_peaks(local, 4 << 30) if rank == 0 else None
    shard_bytes = float(g.numel() * 8)
    eng.close()
    del eng, g
    torch.cuda.empty_cache()

    # ---------------- config 5: M=400, N=24 ------------------------------------------------------
    config5 = None
    if headline and not args.no_config5:
        try:
            M5, N5, K5 = 400, 24, max(5, min(K, 20))
            packed5 = world == 1
            eng5, g5, t05, mloc5, ar5 = build_engine(M5, N5, dev, rank, world, args, packed5)
            eng5.use_stream(stream)
            U5 = [synthetic.random_partial_unitary(M5, N5, seed=synthetic.SEED_U + i).to(dev)
                  for i in range(K5 + 3)]
            ms5, _w5 = timed_evals(eng5, U5, K5, 3, stream, barrier, world, dev)
            k1_5, k1_5_all, parts5 = k1_timing(eng5, U5, K5, 3, world, dev)
            slabs5 = eng5.streamed_slabs()
            b5 = 8.0 * slabs5 * M5 ** 2
            f5 = slabs5 * (2.0 * M5 ** 2 * N5 + 2.0 * M5 * N5 ** 2)
            config5 = {"workload": "synthetic 8-fold-symmetric ERI M=400, N=24 (BASELINE.json "
                                   "configs[4])", "value": K5 / (ms5 * 1e-3), "unit": "evals/s",
                       "steps": K5, "ms_per_step": ms5 / K5,
                       "storage": "pair-packed, one GPU (102.7 GB)" if packed5 else
                       f"dense first-index shard, {mloc5} rows/GPU",
                       "eri_shard_bytes_per_gpu": float(g5.numel() * 8), "allreduce": ar5,
                       "k1_ms_per_launch": k1_5,
                       "k1_tflops": f5 / (k1_5 * 1e-3) / 1e12, "k1_gbs": b5 / (k1_5 * 1e-3) / 1e9,
                       "kernel_ms": {"k1_half_transform": parts5[0], "k_prepare_q": parts5[1],
                                     "k_tail_reduce": parts5[2], "eval_total": parts5[4]}}
            eng5.close()
            del eng5, g5
            torch.cuda.empty_cache()
        except Exception as exc:                  # the headline line must survive
            config5 = {"error": f"{type(exc).__name__}: {exc}"}

    clocks = None
    if rank == 0:
        # one nvidia-smi stream for the whole run, cut by the host time stamps of the regions
        clocks = sampler.stop(window)
        if sustained is not None:
            sustained["clocks"] = _window_clocks(sampler, sustained.pop("window"))
    elif sustained is not None:
        sustained.pop("window", None)
    if rank == 0:
        hbm_peak, peak_src = _load_peaks()
        dmma_cold, dfma_cold, read_cold = cold
        dmma, dfma, stream_read = hot
        # algorithmic work of K1 = the M x M slabs it streams (dense: all mloc*M of the shard;
        # pair-symmetric: one of every pair (t,q)/(q,t), Y[q,t] = Y[t,q]^T)
        alg_bytes = 8.0 * slabs * M ** 2
        alg_flops = slabs * (2.0 * M ** 2 * N + 2.0 * M * N ** 2)
        ach_gbs = alg_bytes / (k1_avg * 1e-3) / 1e9
        ach_tf = alg_flops / (k1_avg * 1e-3) / 1e12
        traffic = _load_traffic(M, N, mloc, slabs)
        if config5 and "k1_tflops" in config5:
            config5["roofline_tensor"] = {"frac": config5["k1_tflops"] / dmma,
                                          "frac_cold": config5["k1_tflops"] / dmma_cold}
            config5["roofline_hbm"] = {"frac": config5["k1_gbs"] / hbm_peak,
                                       "frac_read_peak": config5["k1_gbs"] / read_cold}
        line = {
            "metric": METRIC if headline else
            f"orbital-opt energy+grad evals/sec at M={M},N={N} (FP64)",
            "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if headline else
                       f"synthetic 8-fold-symmetric ERI M={M}, N={N}",
                       "M": M, "N": N,
                       "eri_shard_bytes_per_gpu": shard_bytes,
                       "storage": "pair-packed (streamed slabs only)" if args.packed else
                       "dense first-index shard",
                       "slab_mode": "dense" if args.dense else
                       "pair-symmetric (one slab per (t,q)/(q,t) pair" +
                       (", symmetry verified on device)" if world == 1 and not args.packed else
                        ", symmetric by construction)"),
                       "slabs_streamed_per_eval_per_gpu": slabs,
                       "eri_bytes_streamed_per_eval_per_gpu": alg_bytes,
                       "sharding": f"ERI first index over {world} GPU(s), rows/GPU={mloc}",
                       "allreduce": ar_mode,
                       "timed_region": f"burst of {K} evaluations = {ms_total:.1f} ms; see "
                                       f"`sustained` for >= 2 s of the same loop",
                       "cache": "ERI shard (>=4.3 GB) is larger than the 126 MB L2 and a fresh U "
                                "is used every step; no explicit L2 flush"},
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "frac_read_peak": ach_gbs / read_cold,
                         "read_peak": read_cold,
                         "read_peak_source": "read-only LDG.128 stream over 4 GiB measured live on "
                                             "the idle GPU (oo_measure_peaks); K1 only reads",
                         "kernel": "k1_half_transform", "ms_per_launch": k1_avg,
                         "ms_per_launch_median_rank0": sorted(k1_ms)[len(k1_ms) // 2],
                         "ms_per_launch_min_rank0": min(k1_ms),
                         "algorithmic_bytes_per_launch": alg_bytes},
            "roofline_tensor": {"bound": "tensor", "achieved": ach_tf, "peak": dmma,
                                "unit": "TFLOP/s", "frac": ach_tf / dmma,
                                "peak_cold": dmma_cold, "frac_cold": ach_tf / dmma_cold,
                                "peak_source": "DMMA.8x8x4 register-resident loop measured live "
                                               "(oo_measure_peaks): `peak` after the timed loops "
                                               "(power capped), `peak_cold` on the idle GPU before "
                                               "them; DFMA pipe = %.1f TFLOP/s" % dfma_cold,
                                "algorithmic_flops_per_launch": alg_flops},
            "kernel_ms": {"k1_half_transform": parts[0], "k_prepare_q": parts[1],
                          "k_tail_reduce": parts[2], "eval_total": parts[4]},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": M * N * 8,
                    "d2h_bytes_per_step": (M * N + 1) * 8,
                    "api": "OrbitalEngine.submit_host / wait_host -> oo_eval_submit / oo_eval_wait "
                           "(host buffers, two evaluations in flight)",
                    "sync_value": e2e_sync_value,
                    "sync_api": "OrbitalEngine.energy_grad_host -> oo_energy_grad_host (one "
                                "evaluation in flight)"},
            "sustained": sustained,
            "parity": parity,
            "config5": config5,
            "inner_loop": {"iterations_per_s": inner_rate, "iterations": inner_iters,
                           "newton_schulz_iterations_per_retraction":
                               res["newton_schulz_iterations"] / max(1, inner_iters),
                           "jacobi_fallbacks": res["jacobi_fallbacks"],
                           "what": "oo_optimize: device-resident loop of pupo.py:161-350, one "
                                   "evaluation + retraction + BB step + stop test per iteration "
                                   "(3 launches), host round trip only every 4 iterations"},
            "gpu_launches": int(launches_per_eval * K),
            "gpu_launches_per_eval": int(launches_per_eval),
            "clocks": clocks,
            "energy_checksum": e_sum,
        }
        if not args.no_cpu_baseline and world == 1:     # rank 0 at N=1 only (bench contract)
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
