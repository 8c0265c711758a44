"""bench.py -- collocation points/s per full train step (fwd + residual + bwd + Adam [+ DP sync]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): examples/poisson_2D_1C -- FNN 2->[64]x9->1 (Sine + erf-GELU),
1,000,000 domain points (Laplacian residual, jets of order 2) + 4 x 50,000 boundary points per GPU,
calc_loss="mean", Adam(1e-3), synthetic uniform points, random-init weights (seed 42).  Weak scaling:
every rank holds that many points; `value` = all ranks' points / max-over-ranks step time.

One JSON line on stdout (rank 0).  `--impl reference` times the reference's algorithm (nested
torch.autograd with TorchScript-fused activations, oracle/autograd_ref.py) on the host CPU cores on a bounded
sample (100 000 domain points).
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
def reference_model_and_points(workload, n_domain, n_bound, device, state_dict=None, points=None):
    from oracle import autograd_ref as AR

    key = WORKLOADS[workload][0]
    fn, kind, kw = problems.CONFIGS[key]
    ns = types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack, Symbols=None)
    spec, din, dout = fn(ns)
    torch.manual_seed(SEED)
    model = AR.MODELS[kind](din, dout, **kw).to(device)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    if points is None:
        rng = np.random.default_rng(SEED)
        points = sample_all(spec, n_domain, n_bound, rng)
    pts = [torch.as_tensor(p, device=device).detach().clone().requires_grad_() for p in points]
    return AR, model, [f for f, _ in spec], pts


def reference_cpu_run(workload, steps, warmup, n_domain, n_bound, device="cpu", scripted=True):
    """points/s of the reference's nested-autograd train step (oracle/autograd_ref.py) on `device`."""
    AR, model, fns, pts = reference_model_and_points(workload, n_domain, n_bound, device)
    AR.SCRIPTED = scripted
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
    AR.SCRIPTED = True
    return n_points / dt, dt * 1e3, n_points, float(total)


def cpu_sample_sizes(workload):
    """Bounded CPU sample: BASELINE.md section 3.3's 100 000 domain points (+ 5 000 per boundary condition)."""
    if workload == "regression_1D":
        return 256, 256
    return 100_000, 5_000


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nd, nb = cpu_sample_sizes(args.workload)
    value, ms, n_points, _ = reference_cpu_run(args.workload, args.steps, args.warmup, nd, nb)
    sample = (f"{nd} domain + {nb} per boundary condition points per step ({n_points} total), fp32, torch CPU, "
              f"{args.steps} timed steps after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": "collocation points/sec per full train step",
        "value": value, "unit": "points/s", "n_gpus": 1, "requested_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample,
                   "algorithm": "nested torch.autograd.grad + backward + Adam, TorchScript-fused activations as in the "
                                "reference (oracle/autograd_ref.py port); ONE CPU process whatever --gpus says"},
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


def ncu_traffic(workload, n_domain, kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture of this command
    (profiles/r2_ncu_traffic.json, written by scripts/ncu_traffic.py from one `ncu --set full` run); None if the
    capture does not cover this workload / size."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        ent = tab.get(f"{workload}:{n_domain}", {}).get(kernel)
        return ent
    except Exception:
        return None


def run_ours(args):
    import torch.distributed as dist

    import multipinn_b200 as mp
    from multipinn_b200 import _lib, engine, trace

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
    if args.no_laplace:
        trace.ENABLE_LAPLACE = False

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
    host_np = sample_all(spec, n_domain, n_bound, rng)
    host_pts = [torch.from_numpy(p).pin_memory() for p in host_np]
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
    trainer.use_fused_step = not args.no_fused
    trainer.use_cuda_graphs = not args.no_graphs
    pinn.update_data()
    local_points = sum(len(p) for p in dev_pts)
    global_points = local_points * world
    fs = trainer.fused_step()
    programs = [c.program_for(model) for c in conds]
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}

    # ---------------- step-0 loss and flat gradient of this arm (compared with the reference algorithm below)
    ours0 = None
    if world == 1 and rank == 0 and not args.no_baselines:
        pinn.select_batch(0)
        if fs is not None:
            fs.compute()
            tot0, _ = fs.losses()
        else:
            model.jet_engine.zero_grad()
            tot0, _ = pinn.calculate_mean_loss()
            tot0.backward()
        ours0 = (float(tot0), model.jet_engine.flat_grad.detach().clone())

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
    ms = e0.elapsed_time(e1) / args.steps
    clock_info = clocks.stop() if rank == 0 else None
    final_loss = float(trainer.total_loss)
    graph_replays = fs.graph_replays if fs is not None else 0

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

    # The losses of every step are copied to pinned host memory asynchronously and READ one step later (what a
    # training loop that logs its loss does without stalling the device): the host reads step k's numbers while
    # step k+1 runs, so the device never waits for the host to prepare the next launch.  The last step's losses are
    # read inside the timed region too.
    K1 = 1 + int(trainer.epoch_loss_detailed.numel())
    host_loss = [torch.empty(K1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]   # step k finished reading bufs[k]
    loss_log = []

    def read_losses(k):
        loss_ready[k].synchronize()
        loss_log.append((float(host_loss[k][0]), host_loss[k][1:].clone()))

    def e2e_step(epoch, first=False):
        k = epoch & 1
        torch.cuda.current_stream().wait_event(ready[k])
        for i in range(len(host_pts)):
            dev_pts[i] = bufs[k][i].detach().requires_grad_()
        pinn.update_data()
        step(epoch)
        consumed[k].record()
        host_loss[k][0:1].copy_(trainer.total_loss.detach().reshape(1), non_blocking=True)
        host_loss[k][1:].copy_(trainer.epoch_loss_detailed.detach().reshape(-1), non_blocking=True)
        loss_ready[k].record()
        if not first:
            read_losses(k ^ 1)            # the previous step's losses (its compute is done or about to be)
        copy_stream.wait_event(consumed[k ^ 1]) if not first else None
        prefetch(k ^ 1)                   # next step's points, into the buffer the previous step has finished with

    prefetch(epoch & 1)
    e2e_step(epoch, first=True)
    epoch += 1
    torch.cuda.synchronize()
    loss_log.clear()
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step(epoch)
        epoch += 1
    read_losses((epoch - 1) & 1)          # the last step's losses
    f1.record()
    barrier()
    assert len(loss_log) == args.steps + 1 and all(l == l for l, _ in loss_log)
    e2e_ms = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3) / args.steps
    h2d = sum(p.numel() * 4 for p in host_pts)
    d2h = 4 * (1 + int(trainer.epoch_loss_detailed.numel()))

    # ---------------- attribution pass: the same step launched kernel by kernel (no graph) with CUDA events
    # around every launch of ours, to get the dominant kernel's average duration
    trainer.use_cuda_graphs = False
    engine.CudaBackend.timings = []
    for _ in range(3):
        step(epoch)
        epoch += 1
    torch.cuda.synchronize()
    timings = engine.CudaBackend.timings
    engine.CudaBackend.timings = None
    trainer.use_cuda_graphs = not args.no_graphs

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    replicas_in_sync = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # data-parallel correctness on NCCL: after the same number of steps every rank must hold the same weights
        flat = model.jet_engine.flat.detach().double()
        mine = torch.stack([flat.sum(), flat.abs().sum(), (flat * flat).sum()])
        allsums = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
        replicas_in_sync = all(bool(torch.equal(a, allsums[0])) for a in allsums)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        # dominant kernel: the reverse pass over the domain condition
        Pw = weight_elems(model)
        J_ref = 1 + 2 * din          # channels of the reference formulation ("J used")
        dom_orders = ([b.order for b in programs[0].blocks] if programs[0] is not None
                      else conds[0].plan_for(model).orders)
        order_dom = max(dom_orders)
        J_exec = model.jet_engine.n_jets(order_dom)
        per_kind = {}
        for a in timings:
            per_kind.setdefault((a[0], a[1], a[2]), []).append(a[3].elapsed_time(a[4]))
        avg = {k: sum(v) / len(v) for k, v in per_kind.items()}
        bwd_ms = avg.get(("bwd", n_domain, order_dom))
        fwd_ms = avg.get(("fwd", n_domain, order_dom))
        res_ms = avg.get(("resid", n_domain, 0))
        kernel_ms_per_step = sum(sum(v) for v in per_kind.values()) / 3.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained bf16"
        tf32_peak = tf32_matmul_peak()
        roofline = None
        if bwd_ms:
            flops = n_domain * J_ref * 4 * Pw
            hidden = sum(int(r[2]) for r in model.jet_program()[:-1])   # stashed pre-activation features per point
            stash_bytes = n_domain * J_exec * hidden * 4
            hbm_peak = peaks.get("hbm_gbs", 6500.0)
            ach = flops / (bwd_ms * 1e-3) / 1e12
            widths = {int(r[2]) for r in model.jet_program()[:-1]}
            if not model.jet_engine.tc_supported(order_dom):
                kname = "jet_bwd_ffma"
            elif widths == {128}:   # layer-wise engine (csrc/jet_tc_wide.cu): avg_launch_ms is the whole reverse SEQUENCE
                kname = (f"layer-wise tcgen05 sequence: wide_last_bwd, then wide_wgrad_kernel<{din}, {order_dom}> + "
                         f"wide_dgrad_layer_kernel<{din}, {order_dom}> per 128x128 layer, wide_first_bwd, fold")
            else:
                kname = f"jet_bwd_tc<{din}, {order_dom}, *>"
            roofline = {"bound": "tensor", "kernel": f"jet reverse pass, domain condition ({kname})",
                        "achieved": ach, "peak": peak_bf16, "unit": "TFLOP/s", "frac": ach / peak_bf16,
                        "peak_source": peak_src + " (bf16 dense; the TF32 matmul peak measured in this run is "
                                                  "tf32_matmul_peak_measured)",
                        "traffic": ncu_traffic(args.workload, n_domain, "bwd"),
                        "traffic_unit": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum from the "
                                        "committed ncu capture (profiles/r2_ncu_traffic.json); null = not captured",
                        # N * J * 4 * P_w with J of the reference formulation (value, d first, d second
                        # derivatives); in Laplace mode the kernels carry J_exec channels for the same result
                        "algorithmic_flops_per_launch": flops, "channels_reference": J_ref,
                        "channels_executed": J_exec, "executed_flops_per_launch": n_domain * J_exec * 4 * Pw,
                        "avg_launch_ms": bwd_ms,
                        "tf32_matmul_peak_measured": tf32_peak,
                        "frac_of_tf32_peak": (ach / tf32_peak) if tf32_peak else None,
                        "stash_bytes_per_launch": stash_bytes, "hbm_peak_GBps": hbm_peak,
                        "hbm_frac_reverse": stash_bytes / (bwd_ms * 1e-3) / 1e9 / hbm_peak,
                        "hbm_frac_forward": (stash_bytes / (fwd_ms * 1e-3) / 1e9 / hbm_peak) if fwd_ms else None,
                        "fwd_avg_launch_ms": fwd_ms, "residual_program_avg_launch_ms": res_ms,
                        "fwd_achieved": (n_domain * J_ref * 2 * Pw / (fwd_ms * 1e-3) / 1e12) if fwd_ms else None,
                        "fwd_traffic": ncu_traffic(args.workload, n_domain, "fwd"),
                        "our_kernels_ms_per_step_unfused_launches": kernel_ms_per_step,
                        # the ALGORITHM needs only the points twice and the parameters / gradients;
                        # the pre-activation stash between the forward and the reverse kernel is this design's
                        # traffic, ~1000x that (stash_bytes_per_launch written once and read once)
                        "algorithmic_hbm_bytes_per_step": sum(len(p) for p in dev_pts) * din * 4 * 2
                        + 3 * sum(p.numel() for p in model.parameters()) * 4}
            if "layer-wise" in kname:
                # layer-wise engine: every layer pass reads / writes whole plane sets (J_exec x 128 x 4 B per point);
                # forward 1 + 2 per square layer + 1, reverse 2 + 5 per square layer + 1
                n_sq = len(model.jet_program()) - 2
                plane = n_domain * J_exec * 128 * 4
                roofline["design_plane_bytes_forward"] = (2 + 2 * n_sq) * plane
                roofline["design_plane_bytes_reverse"] = (3 + 5 * n_sq) * plane
                roofline["hbm_frac_forward"] = ((2 + 2 * n_sq) * plane / (fwd_ms * 1e-3) / 1e9 / hbm_peak) if fwd_ms else None
                roofline["hbm_frac_reverse"] = (3 + 5 * n_sq) * plane / (bwd_ms * 1e-3) / 1e9 / hbm_peak
                roofline["note"] = ("HBM-bound by design: hbm_frac_* = plane bytes the layer-wise kernels must move "
                                    "/ time / measured HBM peak; frac (tensor) is against the algorithmic flops")
            tr_b, tr_f = roofline["traffic"], roofline["fwd_traffic"]
            roofline["measured_dram_bytes_domain_fwd_plus_bwd"] = (tr_b + tr_f) if (tr_b and tr_f) else None
        step_flops = 0
        for i in range(len(spec)):
            if programs[i] is not None:
                orders = [min(b.order, 2) for b in programs[i].blocks]
            else:
                orders = conds[i].plan_for(model).orders
            for order in orders:
                step_flops += len(dev_pts[i]) * (1 + order * din) * 6 * Pw
        line = {
            "metric": "collocation points/sec per full train step", "value": global_points / (ms * 1e-3),
            "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tensor-core products as 3xTF32 / BF16x3 splits, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {kind} {kw}, {n_domain} domain + {len(spec) - 1}x{n_bound} "
                                   f"boundary points per GPU, calc_loss=mean, Adam(1e-3)+ExponentialLR",
                       "engine": args.engine, "points_per_gpu": local_points, "regularization": args.regularization,
                       "step_path": ("fused step (traced residual programs)" if fs is not None else
                                     "eager jet path (autograd + torch element-wise residuals)"),
                       "cuda_graphs": bool(fs is not None and not args.no_graphs), "graph_replays": graph_replays,
                       "traced_conditions": sum(p is not None for p in programs), "conditions": len(programs),
                       "untraced_reasons": [c.trace_note for c in conds if c.trace_note],
                       "domain_jet_order": order_dom, "domain_channels": J_exec,
                       "laplace_mode": order_dom == 3,
                       "l2": "per-step working set (pre-activation stash, GBs) exceeds L2; no explicit flush",
                       "algorithmic_tflop_per_step_per_gpu": step_flops / 1e12},
            "clocks": clock_info,
            "e2e": {"value": global_points / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": roofline,
            "final_loss": final_loss,
        }
        if replicas_in_sync is not None:
            line["replicas_in_sync"] = replicas_in_sync
        if world == 1 and not args.no_baselines:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            nd, nb = cpu_sample_sizes(args.workload)
            c_steps, c_warm = 6, 2
            v, cms, npts, _ = reference_cpu_run(args.workload, c_steps, c_warm, nd, nb)
            line["cpu_baseline"] = {"value": v, "unit": "points/s", "cores": torch.get_num_threads(),
                                    "kind": "port", "ms_per_step": cms,
                                    "sample": f"{nd} domain + {nb}/boundary points per step ({npts} total), "
                                              f"{c_steps} timed steps after {c_warm} warm-up, nested torch.autograd "
                                              "with TorchScript-fused activations on CPU"}
            # the reference's own torch-CUDA autograd step on this GPU: the SAME point counts as our arm when
            # its nested-autograd graph fits in memory, else a bounded sample; with the reference's TorchScript
            # activations (its real code path) and, for comparison, with plain eager activations
            sizes = [(256, 256)] if args.workload == "regression_1D" else [(n_domain, n_bound), (100_000, 5_000)]
            for label, scripted in (("ref_torch_cuda", True), ("ref_torch_cuda_eager_activations", False)):
                for nd2, nb2 in sizes:
                    try:
                        torch.cuda.empty_cache()
                        v2, cms2, npts2, _ = reference_cpu_run(args.workload, 3, 3, nd2, nb2, device="cuda",
                                                               scripted=scripted)
                        line[label] = {"value": v2, "unit": "points/s", "ms_per_step": cms2,
                                       "speedup_of_this_arm": line["value"] / v2,
                                       "sample": f"{nd2} domain + {nb2}/boundary points ({npts2} total), nested "
                                                 "torch.autograd fp32 (TF32 off) on the same B200, activations "
                                                 + ("@torch.jit.script as in the reference" if scripted else "eager")}
                        break
                    except Exception as ex:  # out of memory etc.
                        line[label] = {"unavailable": str(ex)[:200]}
            # parity at the bench's full size: step-0 loss and flat parameter gradient, this arm vs the reference
            # algorithm (fp32 and, if it fits, fp64) on the same points and weights
            if ours0 is not None:
                par = {"points": local_points}
                for dt_name, dt in (("fp32", torch.float32), ("fp64", torch.float64)):
                    try:
                        torch.cuda.empty_cache()
                        AR, rm, fns, rpts = reference_model_and_points(
                            args.workload, n_domain, n_bound, "cuda", sd0, host_np)
                        rm = rm.to(dt)
                        rpts = [p.detach().to(dt).requires_grad_() for p in rpts]
                        rt, _ = AR.mean_loss(rm, fns, rpts)
                        rt.backward()
                        rg = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                                        for p in rm.parameters()]).double()
                        og = ours0[1].double()
                        par[f"loss_rel_err_vs_{dt_name}"] = abs(ours0[0] - float(rt)) / abs(float(rt))
                        par[f"grad_rel_err_vs_{dt_name}"] = float((og - rg).abs().max() / rg.abs().max())
                        # the same, normalised inside every parameter tensor (the global maximum sits in the last
                        # layer and would hide the earlier ones): worst tensor
                        worst = 0.0
                        for p_, off_ in zip(model.jet_engine._params, model.jet_engine._offsets):
                            sl_ = slice(off_, off_ + p_.numel())
                            den_ = float(rg[sl_].abs().max())
                            if p_.requires_grad and den_ > 0:
                                worst = max(worst, float((og[sl_] - rg[sl_]).abs().max()) / den_)
                        par[f"grad_rel_err_worst_tensor_vs_{dt_name}"] = worst
                        del rm, rpts, rt, rg
                    except Exception as ex:
                        par[f"{dt_name}_unavailable"] = str(ex)[:120]
                line["parity_full_size"] = par
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
    ap.add_argument("--no-fused", action="store_true", help="A/B: eager jet path (autograd + torch residual algebra)")
    ap.add_argument("--no-graphs", action="store_true", help="A/B: fused step launched kernel by kernel")
    ap.add_argument("--no-laplace", action="store_true", help="A/B: keep d second-order channels (no Laplace mode)")
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
