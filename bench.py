"""bench.py -- collocation points/s per full train step (fwd + residual + bwd + Adam [+ DP sync]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): examples/poisson_2D_1C -- FNN 2->[64]x9->1 (Sine + erf-GELU),
1,000,000 domain points (Laplacian residual, jets of order 2) + 4 x 50,000 boundary points per GPU,
calc_loss="mean", Adam(1e-3), synthetic uniform points, random-init weights (seed 42).  Weak scaling:
every rank holds that many points; `value` = all ranks' points / max-over-ranks step time.

One JSON line on stdout (rank 0).  `--impl reference` times the reference's algorithm (nested
torch.autograd, oracle/autograd_ref.py) on the host CPU cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import problems  # noqa: E402  (workload definitions shared with the parity tests)

WORKLOADS = {
    # name: (problem key, domain points per GPU, boundary points per condition per GPU)
    "poisson_2D_1C": ("poisson_2D_1C", 1_000_000, 50_000),
    "burgers_1D": ("burgers_1D", 4_000_000, 50_000),
    "allen_cahn_2D_1C": ("allen_cahn_2D_1C", 4_000_000, 50_000),
    "navier_stokes_2D": ("navier_stokes_2D", 1_000_000, 50_000),
    "navier_stokes_3D": ("navier_stokes_3D", 2_000_000, 50_000),
    "regression_1D": ("regression_1D", 256, 256),
}
SEED = 42


def env_int(name, default):
    return int(os.environ.get(name, default))


def sample_all(spec, n_domain, n_bound, rng):
    return [problems.sample_points(lo, hi, n_domain if i == 0 else n_bound, rng)
            for i, (_, (lo, hi)) in enumerate(spec)]


def weight_elems(model):
    return sum(p.numel() for p in model.parameters() if p.dim() == 2 and p.requires_grad)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's algorithm on host cores
# --------------------------------------------------------------------------------------------------
def reference_cpu_run(workload, steps, warmup, n_domain, n_bound, device="cpu"):
    """points/s of the nested-autograd step (oracle/autograd_ref.py) on `device`."""
    from oracle import autograd_ref as AR

    key = WORKLOADS[workload][0]
    fn, kind, kw = problems.CONFIGS[key]
    ns = types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack, Symbols=None)
    spec, din, dout = fn(ns)
    torch.manual_seed(SEED)
    model = AR.MODELS[kind](din, dout, **kw).to(device)
    rng = np.random.default_rng(SEED)
    pts = [torch.tensor(p, device=device).requires_grad_() for p in sample_all(spec, n_domain, n_bound, rng)]
    fns = [f for f, _ in spec]
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    n_points = sum(len(p) for p in pts)
    for _ in range(warmup):
        AR.train_step(model, opt, fns, pts)
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        total, _ = AR.train_step(model, opt, fns, pts)
    if device != "cpu":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return n_points / dt, dt * 1e3, n_points, float(total)


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nd, nb = 20_000, 1_000
    if args.workload == "regression_1D":
        nd = nb = 256
    value, ms, n_points, _ = reference_cpu_run(args.workload, args.steps, args.warmup, nd, nb)
    sample = f"{nd} domain + {nb} per boundary condition points per step ({n_points} total), fp32, torch CPU"
    line = {
        "impl": "reference", "metric": "collocation points/sec per full train step",
        "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample,
                   "algorithm": "nested torch.autograd.grad + backward + Adam (oracle/autograd_ref.py port)"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def tf32_matmul_peak():
    """cuBLAS TF32 8192^3 matmul, best of 5 (TFLOP/s) -- the tensor-pipe denominator for 3xTF32 work."""
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device="cuda")
        b = torch.randn(8192, 8192, device="cuda")
        best = 0.0
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = prev
        return best
    except Exception:
        return None


def run_ours(args):
    import torch.distributed as dist

    import multipinn_b200 as mp
    from multipinn_b200 import _lib, engine

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    key, n_domain, n_bound = WORKLOADS[args.workload]
    if args.domain_points:
        n_domain = args.domain_points
    if args.bound_points:
        n_bound = args.bound_points
    fn, kind, kw = problems.CONFIGS[key]
    ns = types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack, Symbols=mp.Symbols)
    spec, din, dout = fn(ns)
    torch.manual_seed(SEED)
    model = {"FNN": mp.FNN, "ResNet": mp.ResNet, "FFNN": mp.FFNN}[kind](din, dout, **kw).to(dev)
    model.jet_engine.kernel_engine = {"auto": _lib.ENGINE_AUTO, "ffma": _lib.ENGINE_FFMA,
                                      "tcgen05": _lib.ENGINE_TCGEN05}[args.engine]
    rng = np.random.default_rng(SEED + rank)  # only the local shard is generated
    host_pts = [torch.from_numpy(p).pin_memory() for p in sample_all(spec, n_domain, n_bound, rng)]
    dev_pts = [p.to(dev, non_blocking=True).requires_grad_() for p in host_pts]

    class _Gen:  # hands the resident device points to the Condition (no regeneration while timing)
        def __init__(self, i):
            self.i = i
            self.n_points = len(dev_pts[i]) * world

        def generate(self, cond, model):
            return dev_pts[self.i]

    conds = []
    for i, (f, (lo, hi)) in enumerate(spec):
        c = mp.Condition(f, mp.Hypercube(lo, hi))
        c.generator = _Gen(i)
        conds.append(c)
    pinn = mp.PINN(model, conds)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    calc_loss = {"none": lambda: "mean", "gradient": lambda: mp.GradientLosses(alpha=0.9),
                 "normal": lambda: mp.NormalLosses(alpha=0.9),
                 "pointwise": lambda: mp.PointwiseLosses()}[args.regularization]()
    if world > 1:
        trainer = mp.TrainerMultiGPU(pinn, opt, sched, num_epochs=10**9, num_batches=world,
                                     update_grid_every=10**9, calc_loss=calc_loss, rank=rank,
                                     sharded_points=True)  # each rank holds only its own slice
    else:
        trainer = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=10**9, update_grid_every=10**9,
                                     calc_loss=calc_loss)
    pinn.update_data()
    local_points = sum(len(p) for p in dev_pts)
    global_points = local_points * world

    def step(epoch):
        trainer.current_epoch = epoch
        trainer._train_epoch()
        sched.step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi takes a moment to start sampling: launch it before the warm-up, use its output for the
    # timed region (it keeps sampling every 100 ms until stopped)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    epoch = 1
    for _ in range(args.warmup):
        step(epoch)
        epoch += 1

    # ---------------- timed region: K steps, inputs resident in HBM
    barrier()
    engine.CudaBackend.timings = []
    launches0 = engine.CudaBackend.gpu_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(epoch)
        epoch += 1
    e1.record()
    barrier()
    launches = engine.CudaBackend.gpu_launches - launches0
    timings = engine.CudaBackend.timings
    engine.CudaBackend.timings = None
    ms = e0.elapsed_time(e1) / args.steps
    clock_info = clocks.stop() if rank == 0 else None
    final_loss = float(trainer.total_loss)

    # ---------------- e2e: same step through the public API with HOST inputs (pinned) + loss readback
    # Every step's points come from pinned host memory.  The copy of step k+1's points runs on a second stream
    # into the other of two device buffers while step k computes (what a training loop fed from the host does);
    # the step waits on the copy's event before its first kernel, and the losses are read back every step.
    copy_stream = torch.cuda.Stream()
    bufs = [[torch.empty_like(p) for p in dev_pts] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(k):
        with torch.cuda.stream(copy_stream):
            for i, h in enumerate(host_pts):
                bufs[k][i].copy_(h, non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_step(epoch):
        k = epoch & 1
        torch.cuda.current_stream().wait_event(ready[k])
        for i in range(len(host_pts)):
            dev_pts[i] = bufs[k][i].detach().requires_grad_()
        prefetch(k ^ 1)   # (the step that last read that buffer has completed: its losses were read back)
        pinn.update_data()
        step(epoch)
        return float(trainer.total_loss), trainer.epoch_loss_detailed.cpu()

    prefetch(epoch & 1)
    e2e_step(epoch)
    epoch += 1
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step(epoch)
        epoch += 1
    f1.record()
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3) / args.steps
    h2d = sum(p.numel() * 4 for p in host_pts)
    d2h = 4 * (1 + int(trainer.epoch_loss_detailed.numel()))

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        # dominant kernel: the reverse pass over the domain condition (order-2 jets)
        Pw = weight_elems(model)
        J = 1 + 2 * din
        bwd = [a[3].elapsed_time(a[4]) for a in timings if a[0] == "bwd" and a[1] == n_domain and a[2] == 2]
        fwd = [a[3].elapsed_time(a[4]) for a in timings if a[0] == "fwd" and a[1] == n_domain and a[2] == 2]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained bf16"
        tf32_peak = tf32_matmul_peak()
        roofline = None
        if bwd:
            bwd_ms = sum(bwd) / len(bwd)
            flops = n_domain * J * 4 * Pw
            hidden = sum(int(r[2]) for r in model.jet_program()[:-1])   # stashed pre-activation features per point
            stash_bytes = n_domain * J * hidden * 4
            hbm_peak = peaks.get("hbm_gbs", 6500.0)
            ach = flops / (bwd_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "jet reverse pass, domain condition (order-2 jets)",
                        "achieved": ach, "peak": peak_bf16, "unit": "TFLOP/s", "frac": ach / peak_bf16,
                        "peak_source": peak_src,
                        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full`
                        # capture of this exact workload (profiles/r1_tc_ncu_summary.txt); null elsewhere
                        "traffic": (14.334291e9 + 1.166608e9) if (args.workload == "poisson_2D_1C" and
                                                                   args.engine == "auto" and
                                                                   n_domain == 1_000_000) else None,
                        "traffic_unit": "bytes per launch",
                        "algorithmic_flops_per_launch": flops, "avg_launch_ms": bwd_ms,
                        "tf32_matmul_peak_measured": tf32_peak,
                        "frac_of_tf32_peak": (ach / tf32_peak) if tf32_peak else None,
                        # memory view of the same two kernels: the pre-activation jet stash is the algorithmic
                        # HBM traffic (written once by the forward pass, read once by the reverse pass)
                        "stash_bytes_per_launch": stash_bytes,
                        "hbm_peak_GBps": hbm_peak,
                        "hbm_frac_reverse": stash_bytes / (bwd_ms * 1e-3) / 1e9 / hbm_peak,
                        "hbm_frac_forward": (stash_bytes / (sum(fwd) / len(fwd) * 1e-3) / 1e9 / hbm_peak) if fwd else None,
                        "fwd_avg_launch_ms": (sum(fwd) / len(fwd)) if fwd else None,
                        "fwd_achieved": (n_domain * J * 2 * Pw / (sum(fwd) / len(fwd) * 1e-3) / 1e12) if fwd else None}
        step_flops = 0
        for i, (f, _) in enumerate(spec):
            plan = conds[i].plan_for(model)
            for order in plan.orders:
                step_flops += len(dev_pts[i]) * (1 + order * din) * 6 * Pw
        line = {
            "metric": "collocation points/sec per full train step", "value": global_points / (ms * 1e-3),
            "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tensor-core products as 3xTF32 / BF16x3 splits, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {kind} {kw}, {n_domain} domain + {len(spec) - 1}x{n_bound} "
                                   f"boundary points per GPU, calc_loss=mean, Adam(1e-3)+ExponentialLR",
                       "engine": args.engine, "points_per_gpu": local_points, "regularization": args.regularization,
                       "l2": "per-step working set (pre-activation stash, GBs) exceeds L2; no explicit flush",
                       "algorithmic_tflop_per_step_per_gpu": step_flops / 1e12},
            "clocks": clock_info,
            "e2e": {"value": global_points / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": roofline,
            "final_loss": final_loss,
        }
        if world == 1 and not args.no_baselines:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            nd, nb = (256, 256) if args.workload == "regression_1D" else (20_000, 1_000)
            v, cms, npts, _ = reference_cpu_run(args.workload, 8, 2, nd, nb)
            line["cpu_baseline"] = {"value": v, "unit": "points/s", "cores": torch.get_num_threads(),
                                    "kind": "port", "ms_per_step": cms,
                                    "sample": f"{nd} domain + {nb}/boundary points per step ({npts} total), "
                                              "10 steps, nested torch.autograd on CPU"}
            # the reference's own torch-CUDA autograd step on this GPU: the SAME point counts as our arm when
            # its nested-autograd graph fits in memory, else a bounded sample
            sizes = [(256, 256)] if args.workload == "regression_1D" else [(n_domain, n_bound), (100_000, 5_000)]
            for nd2, nb2 in sizes:
                try:
                    torch.cuda.empty_cache()
                    v2, cms2, npts2, _ = reference_cpu_run(args.workload, 3, 2, nd2, nb2, device="cuda")
                    line["ref_torch_cuda"] = {"value": v2, "unit": "points/s", "ms_per_step": cms2,
                                              "speedup_of_this_arm": line["value"] / v2,
                                              "sample": f"{nd2} domain + {nb2}/boundary points ({npts2} total), "
                                                        "nested torch.autograd fp32 (TF32 off) on the same B200"}
                    break
                except Exception as ex:  # out of memory etc.
                    line["ref_torch_cuda"] = {"unavailable": str(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poisson_2D_1C", choices=sorted(WORKLOADS))
    ap.add_argument("--engine", default="auto", choices=["auto", "ffma", "tcgen05"])
    ap.add_argument("--domain-points", type=int, default=0)
    ap.add_argument("--bound-points", type=int, default=0)
    ap.add_argument("--no-baselines", action="store_true")
    ap.add_argument("--regularization", default="none", choices=["none", "gradient", "normal", "pointwise"],
                    help="calc_loss: 'none' = \"mean\" (the BASELINE metric); 'gradient' = GradientLosses(0.9), "
                         "the shipped default of poisson_2D_1C (secondary number)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
