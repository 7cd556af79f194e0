"""Benchmark of the SP-NeRF ray-rendering hot path (BASELINE.json metric: render_rays fwd+bwd rays/s
at 1/2/4/8 B200, + fraction of the MLP tensor-core and compositing HBM rooflines).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (baseline/_ref, else the oracle port)

Workload at every N: BASELINE config 2 per GPU — a training step (forward + backward, --depth --sem,
3 semantic classes, dense labels) on a JAX_269-shaped synthetic batch of 8192 rays x 64 samples
(weak scaling: 8192 rays per rank).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RAYS_PER_GPU = 8192
N_SAMPLES = 64
# algorithmic work of the point network (in=6, C=3): FLOP per sample point
FLOP_FWD, FLOP_DGRAD, FLOP_WGRAD = 5_264_384, 5_249_024, 5_264_384
# algorithmic bytes per ray of compositing (N=64, C=3, fp32)
BYTES_COMP_FWD, BYTES_COMP_BWD = 3612, 5660


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU while the timed region runs: NVML polled every ~5 ms (the timed
    region of the default run is ~0.1-0.3 s, shorter than one nvidia-smi invocation), nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].strip().isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_row(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flag = lambda bit: "Active" if r & bit else "Not Active"     # noqa: E731
        return [str(self.index), str(float(mhz)), str(self.max_mhz), "",
                flag(n.nvmlClocksThrottleReasonHwSlowdown), flag(n.nvmlClocksThrottleReasonHwThermalSlowdown),
                flag(n.nvmlClocksThrottleReasonSwThermalSlowdown), flag(n.nvmlClocksThrottleReasonSwPowerCap)]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self._nvml_row())
                    time.sleep(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                self.nvml = None
            time.sleep(0.1)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][2]) if self.rows and self.rows[0][2].replace(".", "").isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_args(**kw):
    import spnerf_b200  # noqa: F401
    from spnerf_b200 import config
    return config.make_args(sem=True, num_sem_classes=3, fc_units=512, n_samples=N_SAMPLES, **kw)


def build_model(args, device):
    import torch
    from spnerf_b200.models import load_model
    torch.manual_seed(0)
    model = load_model(args)
    with torch.no_grad():                      # "trained-like" density so the transmittance scan is exercised
        model.sigma_from_xyz[0].bias.fill_(3.0)
        model.sigma_from_xyz[0].weight.mul_(4.0)
    return model.to(device)


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on the box's host cores, bounded sample.
# baseline/_ref/ (staged by __graft_entry__.build() from the unmodified checkout, git-ignored, travels with the
# snapshot) is timed when present: kind "reference"; otherwise the oracle restatement: kind "port".
# --------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
CPU_SAMPLE_RAYS = 1024            # BASELINE config 1's batch (modules/opt.py:35 default)


def load_reference():
    """(models, rendering, metrics) modules of the staged reference, or None.  kornia.losses.ssim is only used by an
    eval metric outside the path (modules/metrics.py:7,210-215) and is stubbed."""
    need = ["modules/rendering.py", "modules/__init__.py", "modules/metrics.py", "models/spnerf.py", "models/__init__.py"]
    if not all(os.path.exists(os.path.join(REF_DIR, p)) for p in need):
        return None
    k, kl = types.ModuleType("kornia"), types.ModuleType("kornia.losses")
    kl.ssim = lambda *a, **kw: None
    k.losses = kl
    sys.modules.setdefault("kornia", k)
    sys.modules.setdefault("kornia.losses", kl)
    for name in [m for m in sys.modules if m.split(".")[0] in ("models", "modules")]:
        del sys.modules[name]
    sys.path.insert(0, REF_DIR)
    try:
        import models as ref_models
        from modules import rendering as ref_rendering, metrics as ref_metrics
    finally:
        sys.path.remove(REF_DIR)
    return ref_models, ref_rendering, ref_metrics


def make_cpu_step(rays):
    """One training step (forward + losses + backward) of BASELINE config 2's model on `rays` synthetic rays, on the
    host CPU.  Returns (step function, forward-only function, kind)."""
    import torch
    from spnerf_b200 import config, synthetic
    batch = synthetic.make_batch(rays, seed=269)
    ref = load_reference()
    if ref is not None:
        ref_models, ref_rendering, ref_metrics = ref
        args = config.make_args(sem=True, num_sem_classes=3, fc_units=512, n_samples=N_SAMPLES)
        torch.manual_seed(0)
        model = ref_models.load_model(args)
        with torch.no_grad():
            model.sigma_from_xyz[0].bias.fill_(3.0)
            model.sigma_from_xyz[0].weight.mul_(4.0)
        models = {"coarse": model}
        colour, depth, sem = (ref_metrics.SNerfLoss(lambda_sc=0.0),
                              ref_metrics.DepthLoss(lambda_ds=1.0, GNLL=False, usealldepth=False),
                              ref_metrics.SemanticLoss(lambda_ss=1.0))

        def forward(mode):
            return ref_rendering.render_rays(models, args, batch["rays"], None, semantics=batch["sems"], mode=mode,
                                             valid_depth=batch["valid_depth"], target_depths=batch["depths"],
                                             target_std=batch["depth_std"])

        def step():
            res = forward("train")
            loss = colour(res, batch["rgbs"])[0] + depth(res, batch["depths"][:, 0], batch["depths"][:, 1],
                                                         target_valid_depth=batch["valid_depth"],
                                                         target_std=batch["depth_std"])[0] + sem(res, batch["sems"])[0]
            model.zero_grad(set_to_none=True)
            loss.backward()
            return float(loss.detach())

        def fwd_only():
            with torch.no_grad():
                forward("test")
        return step, fwd_only, "reference"
    from oracle import spnerf_oracle as O
    cfg = O.make_cfg(sem=True, num_sem_classes=3, fc_units=512, n_samples=N_SAMPLES)
    P = {k: v.requires_grad_(True) for k, v in O.random_parameters(cfg, seed=0, sigma_bias=3.0).items()}
    b, n = rays, cfg.n_samples

    def step():
        res = O.render(P, cfg, batch["rays"], None, batch["sems"], "train", batch["valid_depth"], batch["depths"],
                       batch["depth_std"], O.Draws([torch.rand(b, n)], [torch.randn(b, n)]))
        loss = O.colour_loss(res, batch["rgbs"])[0] + O.depth_loss(
            res, batch["depths"][:, 0], batch["depths"][:, 1], batch["valid_depth"], batch["depth_std"], 1.0, False)[0] \
            + O.semantic_loss(res, batch["sems"], 1.0)[0]
        torch.autograd.grad(loss, list(P.values()))
        return float(loss.detach())

    def fwd_only():
        with torch.no_grad():
            O.render(P, cfg, batch["rays"], None, batch["sems"], "test", None, None, None,
                     O.Draws([torch.rand(b, n)], [torch.randn(b, n)]))
    return step, fwd_only, "port"


def time_cpu_path(rays, steps, warmup, forward_reps=0):
    """Median-of-steps timing of the CPU path.  Returns a dict (rays/s from the median step, min, cores, kind)."""
    import statistics
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, fwd_only, kind = make_cpu_step(rays)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    out = {"kind": kind, "cores": cores, "rays": rays, "steps": steps, "median_s": statistics.median(times),
           "min_s": min(times), "total_s": sum(times), "value": rays / statistics.median(times)}
    if forward_reps:
        fwd_only()
        ft = []
        for _ in range(forward_reps):
            t0 = time.perf_counter()
            fwd_only()
            ft.append(time.perf_counter() - t0)
        out["forward_only_value"] = rays / statistics.median(ft)
    return out


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(5, min(a.steps, 8)), max(1, min(a.warmup, 2))
    r = time_cpu_path(CPU_SAMPLE_RAYS, steps, warmup)
    what = ("the UNMODIFIED reference (baseline/_ref: modules/rendering.py, models/spnerf.py, modules/metrics.py)"
            if r["kind"] == "reference" else "the oracle restatement of the reference (baseline/_ref not staged)")
    line = {
        "impl": "reference", "metric": "render_rays_fwd_bwd_rays_per_s", "value": r["value"], "unit": "rays/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["median_s"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: training step fwd+bwd, --depth --sem C=3, 64 samples, fc 8x512; "
                               f"reference CPU path on a bounded sample of {CPU_SAMPLE_RAYS} rays per step "
                               "(BASELINE config 1's batch)",
                   "rays_per_step": CPU_SAMPLE_RAYS, "statistic": "median step", "min_ms_per_step": r["min_s"] * 1e3},
        "cpu_baseline": {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": r["kind"],
                         "cpu": cpu_model_name(),
                         "sample": f"{steps} steps x {CPU_SAMPLE_RAYS} rays x {N_SAMPLES} samples, render_rays + SNerfLoss + "
                                   f"DepthLoss + SemanticLoss + backward, torch fp32 on {r['cores']} host threads: {what}"},
        "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# this framework
# --------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import spnerf_b200
    from spnerf_b200 import engine as E, parallel, synthetic, train_step
    from spnerf_b200.modules import metrics
    from spnerf_b200.modules.rendering import render_rays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    args = make_args()
    model = build_model(args, dev)
    host_batch = {k: v.pin_memory() for k, v in synthetic.make_batch(RAYS_PER_GPU, seed=269 + rank).items()}
    batch = {k: v.to(dev, non_blocking=True) for k, v in host_batch.items()}
    reduce_fn = parallel.allreduce_mean_ if world > 1 else None

    def step(timer=None):
        return train_step.fused_step(model, args, batch, repack=True, timer=timer, allreduce=reduce_fn)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        _, _, scalars, launches = step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        _, _, scalars, launches = step()
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = world * RAYS_PER_GPU * a.steps / (ms_total / 1e3)
    loss_vals = [float(x) for x in scalars[:3]]

    # ---- per-kernel durations (CUDA events on the launching stream), same steps ----
    kern = {}
    prof_steps = max(3, min(a.steps, 10))
    timers = []
    for _ in range(prof_steps):      # back to back like the timed region: the host stays ahead of the device, so an
        t = train_step.StepTimer(True)      # interval between two events holds the kernels between them and no launch gap
        step(t)
        timers.append(t)
    torch.cuda.synchronize()
    for t in timers:
        for k, v in t.durations_ms().items():
            kern[k] = kern.get(k, 0.0) + v / prof_steps
    points = RAYS_PER_GPU * N_SAMPLES
    flop = {"mlp_fwd": FLOP_FWD * points, "mlp_bwd_data": FLOP_DGRAD * points, "mlp_bwd_weights": FLOP_WGRAD * points}
    tc = {k: flop[k] / (kern[k] * 1e-3) / 1e12 for k in flop}
    dominant = max(flop, key=lambda k: kern[k])
    mlp_ms = sum(kern[k] for k in flop)
    mlp_tflops = sum(flop.values()) / (mlp_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dominant)
    roofline = {"bound": "tensor", "kernel": dominant, "achieved": tc[dominant], "peak": pk["tc_sustained"],
                "unit": "TFLOP/s", "frac": tc[dominant] / pk["tc_sustained"], "traffic": traffic,
                "peak_source": pk["source"] + ": sustained fp16/bf16 dense GEMM (kernel timed inside a long step)",
                "frac_of_burst_peak": tc[dominant] / pk["tc_burst"],
                "mlp_all_kernels": {"achieved": mlp_tflops, "frac": mlp_tflops / pk["tc_sustained"],
                                    "frac_of_burst_peak": mlp_tflops / pk["tc_burst"], "ms": mlp_ms,
                                    "flop_per_step": sum(flop.values())},
                "per_kernel_tflops": tc}

    line = {
        "metric": "render_rays_fwd_bwd_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: SP-NeRF training step (fwd+bwd), JAX_269-shaped batch, "
                               "--depth --sem --num_sem_classes 3 --dense_ss, fc 8x512, 64 samples, n_importance 0",
                   "rays_per_gpu": RAYS_PER_GPU, "global_rays": world * RAYS_PER_GPU, "n_samples": N_SAMPLES,
                   "parallelism": f"ray-sharded dp{world}, one NCCL all-reduce of the flat fp32 gradient per step",
                   "l2": "each step streams ~30 GB of saved activations / gradient tiles through HBM (>> 126 MB L2); no flush needed",
                   "step": "weight repack, sampling, fused MLP fwd (CTA pairs), compositing, losses, adjoints, dgrad, wgrad"
                           + (", gradient all-reduce" if world > 1 else "")},
        "gpu_launches": launches * a.steps,
        "loss": {"color": loss_vals[0], "depth": loss_vals[1], "semantic": loss_vals[2]},
        "kernels_ms": kern,
        "roofline": roofline,
        "clocks": sampler.summary(),
    }

    if rank == 0 and world == 1:
        # ---- compositing against the HBM roofline, on a batch large enough to leave L2 ----
        nr = 262144
        out = torch.rand(nr * N_SAMPLES, 11, device=dev)
        out[:, 3] *= 30
        z = torch.sort(torch.rand(nr, N_SAMPLES, device=dev) * 0.2, -1).values.contiguous()
        g_rgb, g_depth, g_sem = (torch.randn(nr, 3, device=dev), torch.randn(nr, device=dev),
                                 torch.randn(nr, 3, device=dev))
        for _ in range(3):
            w, t_, rgb, raw, depth, sem = E.composite_fwd(out, z, 11, 8, 3)
            E.composite_bwd(out, z, w, t_, raw, 11, 8, 3, g_rgb=g_rgb, g_depth=g_depth, g_sem=g_sem)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 5
        ev[0].record()
        for _ in range(reps):
            w, t_, rgb, raw, depth, sem = E.composite_fwd(out, z, 11, 8, 3)
        ev[1].record()
        for _ in range(reps):
            E.composite_bwd(out, z, w, t_, raw, 11, 8, 3, g_rgb=g_rgb, g_depth=g_depth, g_sem=g_sem)
        ev[2].record()
        torch.cuda.synchronize()
        f_ms, b_ms = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
        tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
        f_gbs, b_gbs = BYTES_COMP_FWD * nr / f_ms / 1e6, BYTES_COMP_BWD * nr / b_ms / 1e6
        line["roofline_compositing"] = {
            "bound": "hbm", "achieved": (BYTES_COMP_FWD + BYTES_COMP_BWD) * nr / (f_ms + b_ms) / 1e6,
            "peak": pk["hbm"], "unit": "GB/s",
            "frac": (BYTES_COMP_FWD + BYTES_COMP_BWD) * nr / (f_ms + b_ms) / 1e6 / pk["hbm"],
            "traffic": (tj.get("composite_fwd", 0) + tj.get("composite_bwd", 0)) or None,
            "algorithmic_bytes": (BYTES_COMP_FWD + BYTES_COMP_BWD) * nr,
            "fwd": {"achieved": f_gbs, "frac": f_gbs / pk["hbm"], "ms": f_ms, "traffic": tj.get("composite_fwd")},
            "bwd": {"achieved": b_gbs, "frac": b_gbs / pk["hbm"], "ms": b_ms, "traffic": tj.get("composite_bwd")},
            "rays": nr, "peak_source": pk["source"] + ": copy bandwidth (kernels timed alone)"}
        del out, z, w, t_

    if True:
        # ---- end to end through the public API, host buffers, H2D + D2H inside the timed region ----
        loss_fn, dl, sl = metrics.SNerfLoss(0.0), metrics.DepthLoss(1.0, usealldepth=False), metrics.SemanticLoss(1.0)

        loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        state = {"k": 0, "pending": False, "last": None}

        def e2e_step():
            model.engine.mark_dirty()    # parameters change every optimiser step: the weight repack belongs to the step
            d = {k: v.to(dev, non_blocking=True) for k, v in host_batch.items()}
            res = render_rays({"coarse": model}, args, d["rays"], None, semantics=d["sems"], mode="train",
                              valid_depth=d["valid_depth"], target_depths=d["depths"], target_std=d["depth_std"])
            loss = loss_fn(res, d["rgbs"])[0] + dl(res, d["depths"][:, 0], d["depths"][:, 1],
                                                   target_valid_depth=d["valid_depth"],
                                                   target_std=d["depth_std"])[0] + sl(res, d["sems"])[0]
            for p in model.parameters():
                p.grad = None
            loss.backward()
            if world > 1:                # the .grad tensors are slices of one flat buffer: one in-place all-reduce
                parallel.allreduce_grads_(model)
            # device -> host read of the step's result: every step's loss is copied to pinned memory and read on
            # the host one step later (the way a training loop logs), so the queue never drains at a step boundary
            k = state["k"]
            loss_host[k].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ev[k].record()
            if state["pending"]:
                loss_ev[k ^ 1].synchronize()
                state["last"] = float(loss_host[k ^ 1])
            state["k"], state["pending"] = k ^ 1, True

        def e2e_drain():                 # the last step's loss
            if state["pending"]:
                loss_ev[state["k"] ^ 1].synchronize()
                state["last"] = float(loss_host[state["k"] ^ 1])
                state["pending"] = False
        for _ in range(3):
            e2e_step()
        e2e_drain()
        barrier()
        n_e2e = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        e2e_drain()
        barrier()
        dt_t = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dt = float(dt_t)
        line["e2e"] = {"value": world * RAYS_PER_GPU * n_e2e / dt, "unit": "rays/s",
                       "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host_batch.values())),
                       "d2h_bytes_per_step": 4, "ms_per_step": dt / n_e2e * 1e3, "last_loss": state["last"],
                       "api": "render_rays + SNerfLoss + DepthLoss + SemanticLoss + loss.backward(), weights repacked "
                              "every step, batch copied from pinned host memory every step, every step's loss read "
                              "on the host (one step behind the device)"}

    if rank == 0 and world == 1:
        # ---- the reference's default batch (1024 rays, modules/opt.py:35): host-bound regime.  Wall clock per step
        # (launch overheads are the point here) of the kernel-by-kernel step, of the same step replayed from one
        # CUDA graph, and of the public API (render_rays + loss classes + backward); inputs on the device ----
        b1 = {k: v.to(dev) for k, v in synthetic.make_batch(1024, seed=11).items()}

        def wall_ms(fn, reps=40):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3

        def api_step_1024():
            model.engine.mark_dirty()
            res = render_rays({"coarse": model}, args, b1["rays"], None, semantics=b1["sems"], mode="train",
                              valid_depth=b1["valid_depth"], target_depths=b1["depths"], target_std=b1["depth_std"])
            loss = loss_fn(res, b1["rgbs"])[0] + dl(res, b1["depths"][:, 0], b1["depths"][:, 1],
                                                    target_valid_depth=b1["valid_depth"],
                                                    target_std=b1["depth_std"])[0] + sl(res, b1["sems"])[0]
            for p in model.parameters():
                p.grad = None
            loss.backward()
        eager = wall_ms(lambda: train_step.fused_step(model, args, b1, repack=True))
        launches_1024 = train_step.fused_step(model, args, b1, repack=True)[3]
        api = wall_ms(api_step_1024)
        gs = train_step.GraphedStep(model, args, b1)
        graphed = wall_ms(gs.replay)
        line["small_batch"] = {"rays": 1024, "fused_step_eager_ms": eager, "fused_step_cuda_graph_ms": graphed,
                               "public_api_ms": api, "launches_per_step": launches_1024,
                               "rays_per_s_cuda_graph": 1024 / (graphed / 1e3),
                               "note": "wall clock per step, inputs on the device; round 1 measured ~34 launches per step"}
        del gs

    # ---- the other BASELINE configurations, through the public API (extra objects; the headline stays C2) ----
    if not a.skip_extra:
        from spnerf_b200 import inference
        other = {}
        # C3: --guidedsample --mapping training step, 16384 rays per GPU (64-sample pass + 128-sample pass)
        args3 = make_args(mapping=True, guidedsample=True, chunk=16384)
        model3 = build_model(args3, dev)
        b3 = {k: v.to(dev) for k, v in synthetic.make_batch(16384, seed=300 + rank).items()}
        loss_fn, dl, sl = metrics.SNerfLoss(0.0), metrics.DepthLoss(1.0, usealldepth=False), metrics.SemanticLoss(1.0)

        def c3_step():
            res = render_rays({"coarse": model3}, args3, b3["rays"], None, semantics=b3["sems"], mode="train",
                              valid_depth=b3["valid_depth"], target_depths=b3["depths"], target_std=b3["depth_std"])
            loss = loss_fn(res, b3["rgbs"])[0] + dl(res, b3["depths"][:, 0], b3["depths"][:, 1],
                                                    target_valid_depth=b3["valid_depth"],
                                                    target_std=b3["depth_std"])[0] + sl(res, b3["sems"])[0]
            for p_ in model3.parameters():
                p_.grad = None
            loss.backward()
            if world > 1:
                parallel.allreduce_grads_(model3)
            return loss
        for _ in range(2):
            c3_step()
        barrier()
        e0.record()
        n3 = 3
        for _ in range(n3):
            c3_step()
        e1.record()
        barrier()
        ms3 = torch.tensor([e0.elapsed_time(e1) / n3], device=dev)
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        flop3 = 16384 * (64 * 5381120 + 128 * 16011264)          # pass 1 forward only + pass 2 fwd+bwd
        other["c3_train_guided_mapping"] = {
            "workload": "BASELINE config 3: training step with --guidedsample --mapping (+ --depth --sem), 16384 rays "
                        "per GPU, render_rays + losses + backward through the public API, inputs on the device",
            "rays_per_s": world * 16384 / (float(ms3) / 1e3), "ms_per_step": float(ms3), "steps": n3,
            "mlp_tflops_lower_bound": flop3 / (float(ms3) * 1e-3) / 1e12,
            "frac_of_sustained_tensor_peak_lower_bound": flop3 / (float(ms3) * 1e-3) / 1e12 / pk["tc_sustained"]}
        del model3, b3
        torch.cuda.empty_cache()
        # C4: full-image inference (2048 x 2048 rays), rows split across the ranks, per-ray outputs gathered on rank 0
        side = 2048
        n_img = side * side
        lo, hi = parallel.shard_bounds(n_img, rank, world)
        img_rays = synthetic.make_batch(hi - lo, seed=400 + rank)
        rays4 = img_rays["rays"].to(dev)
        sems4 = img_rays["sems"].to(dev)
        args.chunk = 262144

        def c4_render(n_local):
            local_ = inference.render_image({"coarse": model}, args, rays4[:n_local], None, semantics=sems4[:n_local])
            out_ = {}
            if world > 1:
                for k_ in sorted(local_):
                    out_[k_] = parallel.gather_rays(local_[k_], n_local * world, dst=0)
            return local_, out_
        c4_render(262144 // world)                   # warm-up on one chunk per job
        barrier()
        e0.record()
        c4_render(hi - lo)
        e1.record()
        barrier()
        ms4 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms4, op=dist.ReduceOp.MAX)
        flop4 = n_img * N_SAMPLES * FLOP_FWD
        other["c4_full_image_inference"] = {
            "workload": f"BASELINE config 4: {side}x{side} rays x 64 samples forward, rows sharded over {world} GPU(s), "
                        "chunks of 262144 rays, per-ray rgb/depth/logits/class + composited albedo/sun/sky gathered on rank 0",
            "rays_per_s": n_img / (float(ms4) / 1e3), "seconds_per_image": float(ms4) / 1e3, "scaling": "strong",
            "mlp_tflops_lower_bound": flop4 / (float(ms4) * 1e-3) / 1e12 / world,
            "frac_of_sustained_tensor_peak_lower_bound": flop4 / (float(ms4) * 1e-3) / 1e12 / world / pk["tc_sustained"],
            "d2h_bytes_per_ray_if_exported": 64}
        line["other_configs"] = other
        del rays4, sems4

    if rank == 0 and world == 1:
        # ---- the reference's CPU path on this box's host cores (bounded sample, ~10 s) ----
        try:
            r = time_cpu_path(CPU_SAMPLE_RAYS, 5, 1, forward_reps=3)
            line["cpu_baseline"] = {"value": r["value"], "forward_only_value": r.get("forward_only_value"), "unit": "rays/s",
                                    "cores": r["cores"], "kind": r["kind"], "cpu": cpu_model_name(),
                                    "sample": f"median of 5 steps x {CPU_SAMPLE_RAYS} rays x 64 samples (BASELINE config 1 "
                                              "batch), render_rays + losses + backward, torch fp32; kind 'reference' = the "
                                              "unmodified reference staged under baseline/_ref, 'port' = the oracle"}
        except Exception as ex:          # pragma: no cover
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-extra", action="store_true", help="only the headline (C2) workload: no C3 / C4 objects")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
