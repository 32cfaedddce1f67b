#!/usr/bin/env python
"""bench.py — orbital-optimisation energy+gradient evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3]): synthetic 8-fold-symmetric ERI, M=256 spatial orbitals, N=16
active orbitals, ensemble-N-representable RDMs, FP64.  One step = one (E, dE/dU) evaluation at a
fresh partial unitary U.  With N GPUs the ERI tensor is sharded by its first index (strong scaling
of the same problem) and each evaluation ends with one NCCL all-reduce of M*N+1 doubles.

Prints ONE JSON line (rank 0).  `value` = evaluations/s with U already in HBM; `e2e` = the same
through the host-buffer entry point (H2D of U, D2H of E and dE/dU inside the timed region);
`roofline` = the dominant kernel (K1, the TMA+DMMA half-transform) against the measured HBM peak,
`roofline_tensor` = the same kernel against the DMMA peak measured live; `cpu_baseline` = the
reference's torch formulation (oracle/torch_port.py) on this box's host cores, bounded sample.
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

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(target_seconds=12.0):
    """The reference's torch formulation on the host cores, on a slab of the last ERI index."""
    import torch
    import esoo_b200  # noqa: F401  (synthetic generators live in the package)
    from esoo_b200 import synthetic
    from oracle import torch_port
    M, N = M_BENCH, N_BENCH
    cores = torch.get_num_threads()
    h = synthetic.h_spatial(M)
    D, G = synthetic.rdms_spatial(N)
    U = synthetic.random_partial_unitary(M, N)
    ms = 32
    g = synthetic.eri_spatial_shard(M, 0, ms).permute(3, 2, 1, 0).contiguous()
    torch_port.time_reference(U, D, G, h, g[..., :4].contiguous(), 0)          # warm-up
    t_eval, t_iter = torch_port.time_reference(U, D, G, h, g, 0, repeats=1)
    reps = max(1, min(20, int(target_seconds / max(t_iter, 1e-3))))
    t_eval, t_iter = torch_port.time_reference(U, D, G, h, g, 0, repeats=reps)
    frac = ms / M
    return {"value": frac / t_eval, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": (f"slab of {ms}/{M} of the last ERI index (all contractions of the chain are "
                       f"linear in it), {reps} repeats of einsum forward + autograd backward, "
                       f"{t_eval:.3f} s each; spatial M^4 tensor = 32x less work than the "
                       f"reference's spin-orbital tensor"),
            "reference_iterations_per_s": frac / t_iter}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    steps, vals = max(1, args.steps), []
    base = None
    for _ in range(min(steps, 3)):            # each step = one bounded sample
        base = cpu_baseline(target_seconds=4.0)
        vals.append(base["value"])
    value = sum(vals) / len(vals)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": 1, "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "M": M_BENCH, "N": N_BENCH},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--M", type=int, default=M_BENCH)
    ap.add_argument("--N", type=int, default=N_BENCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    M, N, K, W = args.M, args.N, args.steps, max(3, args.warmup)
    t0, mloc = esoo_b200.shard_range(M, rank, world)
    h = synthetic.h_spatial(M, device=dev)
    D, G = synthetic.rdms_spatial(N)
    eng = esoo_b200.OrbitalEngine(M, N, device=dev, t0=t0, mloc=mloc)
    if args.packed:
        g = synthetic.eri_spatial_pair_packed(M, t0, mloc, device=dev)
        eng.set_integrals_packed(h, g)         # symmetric by construction (half is not stored)
    else:
        g = synthetic.eri_spatial_shard(M, t0, mloc, device=dev)
        if world == 1:
            eng.set_integrals(h, g)            # verifies the V4 symmetry on the device
        else:
            eng.set_integrals(h, g, assume_v4_symmetric=True)   # symmetric by construction
    if world > 1:
        esoo_b200.attach_nccl(eng)
        if args.allreduce == "fused":
            try:
                esoo_b200.attach_peer_memory(eng)
            except Exception as exc:            # no peer access on this box: NCCL still works
                if rank == 0:
                    print(f"# fused all-reduce unavailable ({exc}); using NCCL", file=sys.stderr)
                args.allreduce = "nccl"
            flag = torch.tensor([1 if args.allreduce == "fused" else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0 and args.allreduce == "fused":
                raise SystemExit("ranks disagree on the all-reduce mode")
    eng.set_rdms(D, G)
    if not args.packed:
        eng.set_pair_symmetry(not args.dense)
    slabs = eng.streamed_slabs()
    stream = torch.cuda.Stream(device=dev)
    eng.use_stream(stream)

    # a fresh U per step (identical on every rank), resident in HBM for `value`,
    # in pinned host memory for `e2e`
    Us = [synthetic.random_partial_unitary(M, N, seed=synthetic.SEED_U + i) for i in range(K + W)]
    U_dev = [u.to(dev) for u in Us]
    U_host = [u.numpy() for u in Us]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput --------------------------------------------
    # nvidia-smi needs ~0.1 s to deliver its first sample: start it before the warm-up steps (same
    # kernels, same load) so that short timed regions are still covered
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(W):
        eng.enqueue_energy_grad(U_dev[i])
    barrier()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for i in range(K):
            eng.enqueue_energy_grad(U_dev[W + i])
        ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    value = K / (ms_total * 1e-3)

    # ---------------- per-kernel timing of the dominant kernel (CUDA events, same stream) -----
    eng.set_timing(True)
    k1_ms, parts = [], [0.0] * 5
    for i in range(K):
        eng.enqueue_energy_grad(U_dev[W + i], allreduce=False)
        t = eng.last_timing()
        k1_ms.append(t[0])
        parts = [a + b for a, b in zip(parts, t)]
    eng.set_timing(False)
    k1_avg = sum(k1_ms) / len(k1_ms)
    kt = torch.tensor([k1_avg], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    k1_avg = float(kt.item())

    # ---------------- end to end through host buffers ------------------------------------------
    for i in range(W):
        eng.energy_grad_host(U_host[i])
    barrier()
    t_a = time.perf_counter()
    e_sum = 0.0
    for i in range(K):
        e, _grad = eng.energy_grad_host(U_host[W + i])
        e_sum += e
    torch.cuda.synchronize()
    t_b = time.perf_counter()
    e2e_t = torch.tensor([t_b - t_a], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = K / float(e2e_t.item())

    # ---------------- whole inner loop on the device (eval + all-reduce + retraction + BB + stop) --
    n_inner = max(10, min(K, 50))
    barrier()
    t_c = time.perf_counter()
    res = eng.optimize(U_host[0], 1e-3, 0.0, n_inner)      # tol=0: runs until iteration > maxiter
    t_d = time.perf_counter()
    inner_t = torch.tensor([t_d - t_c], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(inner_t, op=dist.ReduceOp.MAX)
    inner_iters = res["n_iter"]
    inner_rate = inner_iters / float(inner_t.item())

    if rank == 0:
        hbm_peak, peak_src = _load_peaks()
        dmma, dfma, stream_read = esoo_b200.measure_peaks(local, 4 << 30)
        # algorithmic work of K1 = the M x M slabs it streams (dense: all mloc*M of the shard;
        # pair-symmetric: one of every pair (t,q)/(q,t), Y[q,t] = Y[t,q]^T)
        alg_bytes = 8.0 * slabs * M ** 2
        alg_flops = slabs * (2.0 * M ** 2 * N + 2.0 * M * N ** 2)
        ach_gbs = alg_bytes / (k1_avg * 1e-3) / 1e9
        ach_tf = alg_flops / (k1_avg * 1e-3) / 1e12
        traffic = _load_traffic(M, N, mloc, slabs)
        line = {
            "metric": METRIC if (M, N) == (M_BENCH, N_BENCH) else
            f"orbital-opt energy+grad evals/sec at M={M},N={N} (FP64)",
            "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if (M, N) == (M_BENCH, N_BENCH) else
                       f"synthetic 8-fold-symmetric ERI M={M}, N={N}",
                       "M": M, "N": N,
                       "eri_shard_bytes_per_gpu": float(g.numel() * 8),
                       "storage": "pair-packed (streamed slabs only)" if args.packed else
                       "dense first-index shard",
                       "slab_mode": "dense" if args.dense else
                       "pair-symmetric (one slab per (t,q)/(q,t) pair" +
                       (", symmetry verified on device)" if world == 1 and not args.packed else
                        ", symmetric by construction)"),
                       "slabs_streamed_per_eval_per_gpu": slabs,
                       "eri_bytes_streamed_per_eval_per_gpu": alg_bytes,
                       "sharding": f"ERI first index over {world} GPU(s), rows/GPU={mloc}",
                       "allreduce": "none (1 GPU)" if world == 1 else
                       ("one-shot all-reduce fused into k_tail_row over NVLink peer memory"
                        if args.allreduce == "fused" else "NCCL all-reduce of M*N+1 doubles"),
                       "cache": "ERI shard (>=4.3 GB) is larger than the 126 MB L2 and a fresh U "
                                "is used every step; no explicit L2 flush"},
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k1_half_transform", "ms_per_launch": k1_avg,
                         "ms_per_launch_median_rank0": sorted(k1_ms)[len(k1_ms) // 2],
                         "ms_per_launch_min_rank0": min(k1_ms),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "stream_read_gbs_measured_live": stream_read},
            "roofline_tensor": {"bound": "tensor", "achieved": ach_tf, "peak": dmma,
                                "unit": "TFLOP/s", "frac": ach_tf / dmma,
                                "peak_source": "DMMA.8x8x4 register-resident loop measured live "
                                               "(oo_measure_peaks); DFMA pipe = %.1f TFLOP/s" % dfma,
                                "algorithmic_flops_per_launch": alg_flops},
            "kernel_ms": {"k1_half_transform": parts[0] / K, "k_qcontract": parts[1] / K,
                          "k_tail_row": parts[2] / K, "eval_total": parts[4] / K},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": M * N * 8,
                    "d2h_bytes_per_step": (M * N + 1) * 8,
                    "api": "OrbitalEngine.energy_grad_host -> oo_energy_grad_host (host buffers)"},
            "inner_loop": {"iterations_per_s": inner_rate, "iterations": inner_iters,
                           "newton_schulz_iterations_per_retraction":
                               res["newton_schulz_iterations"] / max(1, inner_iters),
                           "jacobi_fallbacks": res["jacobi_fallbacks"],
                           "what": "oo_optimize: device-resident loop of pupo.py:161-350, one "
                                   "evaluation + retraction + BB step + stop test per iteration, "
                                   "host round trip only every 4 iterations (flag poll)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "energy_checksum": e_sum,
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
