#!/usr/bin/env python
"""bench.py — headline measurement of the NeRF volume-rendering hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

metric  : rays/s of one training step (BASELINE.json cfg 2): N_rand = 4096 rays per GPU, 64 coarse +
          (64+64) fine samples, coarse+fine 8x256 MLP fwd+bwd, sample_pdf + raw2outputs, MSE loss on
          rgb and rgb0, gradient allreduce (N>1), Adam step.  Weak scaling: 4096 rays per GPU.
value   : whole-job rays/s with the ray batch already resident in HBM.
e2e     : the same step through the public API (mvip_nerf_b200.run.render) with HOST buffers: rays + targets
          copied from pinned memory every step, loss read back every step.
render  : secondary number — rendered rays/s of BASELINE.json cfg 3 (1008x756 image, render kwargs),
          rays sharded over the GPUs with one final gather.
roofline: the kernel with the largest share of the step, timed with CUDA events inside the timed region.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H, W, FOCAL, NEAR, FAR = 756, 1008, 767.2935, 1.2, 7.7369
N_RAND = 4096
METRIC = "rays/sec (train fwd+bwd, 64+64 samples)"
FLOP_FWD, FLOP_DGRAD, FLOP_WGRAD = 1186816, 1114112, 1185536       # per point, DESIGN.md §kernels
# HBM bytes per point: stash 40 chunk images x 128 B + 288 B masks; dZ stash 38 x 128 B (+ masks read, + 16 B d_raw);
# wgrad reads both sets of images; head grads read hidden (2) + h8 (4) images + d_raw
HBM_FWD_TRAIN, HBM_DGRAD, HBM_WGRAD, HBM_HEADS = 40 * 128 + 288 + 16, 38 * 128 + 288 + 16, 78 * 128, 6 * 128 + 16
# measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch, mean of the coarse (262,144 points) and the
# fine (524,288 points) launch of one step, from the ncu --set full capture committed as profiles/r01_kernels_ncu.md
NCU_TRAFFIC = {"wgrad_kernel": (2.713 + 0.009 + 5.430 + 0.040) / 2 * 1e9,
               "mvip_mlp_forward": (0.018 + 1.389 + 0.033 + 2.834) / 2 * 1e9,
               "dgrad_chain_kernel": (0.087 + 1.216 + 0.175 + 2.491) / 2 * 1e9}


def peaks():
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        d["_source"] = "measured"
        return d
    except Exception:
        return fallback


def nerf_args(basedir):
    return argparse.Namespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=64, N_samples=64, netdepth=8,
        netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, alpha_model_path=None, no_coarse=False,
        lrate=5e-4, basedir=basedir, expname="bench", ft_path=None, no_reload=True, perturb=1.0, white_bkgd=True,
        raw_noise_std=1.0, dataset_type="llff", no_ndc=True, lindisp=True, sigma_loss=False)


def synth_rays_np(n, seed):
    """cfg-2 rays: seeded random pixels of the 1008x756 pinhole grid, identity pose."""
    rng = np.random.RandomState(seed)
    idx = rng.randint(0, H * W, size=n)
    i = (idx % W).astype(np.float32)
    j = (idx // W).astype(np.float32)
    d = np.stack([(i - W * .5) / FOCAL, -(j - H * .5) / FOCAL, -np.ones_like(i)], -1).astype(np.float32)
    o = np.zeros_like(d)
    target = rng.rand(n, 3).astype(np.float32)
    return o, d, target


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """start of the timed region: samples before this instant are ignored"""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t_end = time.time()
        t0 = getattr(self, "t0", 0.0)
        for ts, ln in self.lines:
            if ts < t0 or ts > t_end:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =====================================================================================================
# reference arm: CPU port of the reference path on the host cores
# =====================================================================================================
def cpu_train_rays_per_s(n_rays, steps, warmup):
    from oracle import nerf_oracle as orc
    from oracle import torch_cpu_port as port
    pc = port.params_from_numpy(orc.init_params(1), True)
    pf = port.params_from_numpy(orc.init_params(2), True)
    o, d, target = synth_rays_np(n_rays, 1)
    rays = torch.from_numpy(orc.make_ray_batch(o, d, NEAR, FAR))
    tv = torch.linspace(0., 1., 64)
    g = torch.Generator().manual_seed(0)
    args = (torch.rand(n_rays, 64, generator=g), torch.rand(n_rays, 64, generator=g),
            torch.randn(n_rays, 64, generator=g), torch.randn(n_rays, 128, generator=g), torch.from_numpy(target))
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=5e-4)
    for _ in range(warmup):
        port.train_step(rays, pc, pf, tv, *args); opt.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        port.train_step(rays, pc, pf, tv, *args); opt.step()
    dt = time.perf_counter() - t0
    return n_rays * steps / dt, dt / steps


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 512
    steps = max(1, min(a.steps, 6))
    warm = max(1, min(a.warmup, 1))
    val, sec = cpu_train_rays_per_s(sample, steps, warm)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "cfg2 train step (N_rand=4096/GPU), timed on a %d-ray sample per step" % sample},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": "%d rays x %d steps, PyTorch-CPU port of the reference path (oracle/torch_cpu_port.py)"
                                       % (sample, steps)},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def hbm_stage_times(ops, dev, hbm_peak):
    """Sampling / compositing / ray / normal-map kernels on a 262,144-ray batch (a third of a cfg-3 image): achieved GB/s =
    ALGORITHMIC bytes per launch / CUDA-event time of the launch, best and median of 10."""
    N = 262144
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *sh: torch.rand(*sh, device=dev, generator=g)  # noqa: E731

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(np.min(ts))
    out = {}

    def add(name, nbytes, fn):
        med, best = timeit(fn)
        out[name] = {"bytes_per_launch": nbytes, "ms_median": med, "ms_best": best, "gbs": nbytes / med / 1e6,
                     "frac_of_hbm_peak": nbytes / med / 1e6 / hbm_peak}
    for S in (64, 128):
        raw = rnd(N, S, 4) * 2 - 1
        z = torch.sort(rnd(N, S) * 6 + 1.2, -1)[0]
        rd = rnd(N, 3) - .5
        add("composite_fwd_S%d" % S, N * (S * 20 + 12 + S * 4 + 24), lambda: ops.composite_forward(raw, z, rd, None, True))
        gs = [rnd(N, 3), rnd(N), rnd(N), rnd(N)]
        add("composite_bwd_S%d" % S, N * (S * 20 + 12 + 24 + S * 16), lambda: ops.composite_backward(raw, z, rd, None, True, False, *gs))
    z = torch.sort(rnd(N, 64) * 6 + 1.2, -1)[0]
    w = rnd(N, 64) ** 4
    u = rnd(N, 64)
    add("sample_fine_rand_u", N * (512 + 256 + 512 + 4), lambda: ops.sample_fine(z, w, u, want_samples=False))
    rays = rnd(N, 11) + 1
    tv = torch.linspace(0, 1, 64, device=dev)
    tr = rnd(N, 64)
    add("sample_coarse_perturb", N * (8 + 256 + 256), lambda: ops.sample_coarse(rays, tv, tr, True))
    c2w = torch.eye(4, device=dev)[:3, :4].contiguous()
    add("rays_from_pose_1008x756", H * W * 44, lambda: ops.rays_from_pose(H, W, FOCAL, c2w, NEAR, FAR))
    d = rnd(512, 512) + 3
    add("normal_fwd_512x512_k31", 512 * 512 * 24, lambda: ops.normal_forward(d, 500., 500., 256., 256., 31))
    return out


# =====================================================================================================
# our arm
# =====================================================================================================
def run_ours(a):
    global N_RAND
    N_RAND = int(a.rays_per_gpu)
    from mvip_nerf_b200 import dist as md
    from mvip_nerf_b200 import ops, run
    from mvip_nerf_b200.run_nerf_helpers import img2mse
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local = md.init_from_env()
    dev = torch.device("cuda", local)
    pk = peaks()

    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "bench"))
        torch.manual_seed(0)
        kw_train, kw_test, _, grad_vars, optimizer = run.create_nerf(nerf_args(td))   # optimizer: FusedAdam (one launch)
    coarse, fine = kw_train["network_fn"], kw_train["network_fine"]
    groups = [list(fine.parameters()), list(coarse.parameters())]

    o, d, target = synth_rays_np(N_RAND, 100 + rank)
    pin = lambda x: torch.from_numpy(x).pin_memory()  # noqa: E731
    h_o, h_d, h_t = pin(o), pin(d), pin(target)
    d_rays = torch.stack([h_o.to(dev), h_d.to(dev)], 0)
    d_target = h_t.to(dev)
    inv_world = 1.0 / world

    def step(rays, tgt):
        optimizer.zero_grad(set_to_none=True)
        rgb, disp, acc, depth, extras = run.render(H, W, FOCAL, chunk=32768, rays=rays, near=NEAR, far=FAR, **kw_train)
        loss = (img2mse(rgb, tgt) + img2mse(extras["rgb0"], tgt)) * inv_world     # mean over the GLOBAL batch
        loss.backward()
        md.allreduce_grads(groups)
        optimizer.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- the step as the product runs it: two CUDA graphs per step (graph.GraphedTrainStep); eager fallback if capture fails
    h_rays = torch.stack([h_o, h_d], 0).pin_memory()
    gstep, graph_note = None, "eager launches"
    if not a.no_graph:
        try:
            from mvip_nerf_b200.graph import GraphedTrainStep
            gstep = GraphedTrainStep(kw_train, optimizer, H, W, FOCAL, N_RAND, NEAR, FAR)
            gstep.rays.copy_(d_rays)
            gstep.target.copy_(d_target)
            gstep.capture()
            graph_note = "CUDA graphs: (render + loss + backward) | eager NCCL allreduce | (Adam + bf16 re-pack)"
        except Exception as e:      # noqa: BLE001
            gstep, graph_note = None, "eager launches (graph capture failed: %s)" % str(e)[:120]

    def resident_step():
        return gstep() if gstep is not None else step(d_rays, d_target)

    # ---- device-resident arm ---------------------------------------------------------------------
    for _ in range(a.warmup):
        resident_step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(3):                         # nvidia-smi needs ~100 ms to start streaming: keep the GPUs busy meanwhile
        resident_step()                        # (all ranks: the step contains a collective)
    torch.cuda.synchronize()
    clocks.mark()
    ms_total = timed(resident_step, a.steps)
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / a.steps
    value = world * N_RAND / (ms_step * 1e-3)

    # ---- end-to-end arm: host buffers, H2D every step, loss read back every step -----------------------
    def e2e_step():
        if gstep is not None:
            return float(gstep(h_rays, h_t).item())
        rays = torch.stack([h_o.to(dev, non_blocking=True), h_d.to(dev, non_blocking=True)], 0)
        tgt = h_t.to(dev, non_blocking=True)
        return float(step(rays, tgt).item())
    for _ in range(max(1, a.warmup // 2)):
        e2e_step()
    ms_e2e = timed(e2e_step, a.steps) / a.steps
    e2e_value = world * N_RAND / (ms_e2e * 1e-3)

    # ---- the same K steps launched eagerly with a CUDA-event bracket around every library call: per-kernel times --------
    for _ in range(3):
        step(d_rays, d_target)
    ops.kernel_timer.enable(True)
    l0 = ops.launch_count
    ms_eager = timed(lambda: step(d_rays, d_target), a.steps) / a.steps
    launches = ops.launch_count - l0
    ktimes = ops.kernel_timer.collect()
    ops.kernel_timer.enable(False)

    # ---- secondary: full-image render (cfg 3), rays sharded, one gather ------------------------------
    n_img = H * W
    c2w = torch.eye(4, device=dev)[:3, :4]
    from mvip_nerf_b200.run_nerf_helpers import get_rays
    ro, rd = get_rays(H, W, FOCAL, c2w)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    ones = torch.ones(n_img, 1, device=dev)
    rays_flat = torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3), NEAR * ones, FAR * ones, vd.reshape(-1, 3)], -1).contiguous()
    kw_r = {k: v for k, v in kw_test.items() if k not in ("ndc", "use_viewdirs")}

    def render_image():
        with torch.no_grad():
            return md.render_sharded(lambda rows: run.batchify_rays(rows, 1 << 17, **kw_r), rays_flat)
    render_image()
    r_steps = 3
    ms_render = timed(render_image, r_steps) / r_steps
    render_value = n_img / (ms_render * 1e-3)

    # ---- secondary: guidance batch (cfg 5): 4 views x 512 x 512, rgb + disp + acc + depth, normal maps on rank 0 --------
    GV, GH, GW = 4, 512, 512
    gfocal = FOCAL * GW / W
    gposes = []
    for v in range(GV):
        pz = torch.eye(4, device=dev)[:3, :4].clone()
        pz[0, 3] = 0.1 * v
        gposes.append(pz)

    def render_nograd(*args, **kw):
        with torch.no_grad():
            return run.render(*args, chunk=1 << 17, **kw)

    def guidance():
        return md.render_views_sharded(render_nograd, gposes, GH, GW, gfocal, NEAR, FAR, **kw_test)
    guidance()
    g_steps = 2
    ms_guid = timed(guidance, g_steps) / g_steps
    guid_value = GV * GH * GW / (ms_guid * 1e-3)

    # ---- secondary: one guidance view WITH gradients at full resolution (deferred back-propagation) --------
    # every rank renders its own 512 x 512 view (render kwargs), image loss on rgb + depth, backward through render_deferred
    def guidance_train():
        optimizer.zero_grad(set_to_none=True)
        rgb, disp, acc, depth, _ = run.render_deferred(GH, GW, gfocal, chunk=8192, c2w=gposes[rank % GV], near=NEAR, far=FAR, **kw_test)
        loss = ((rgb - 0.5) ** 2).mean() + 0.1 * depth.mean()
        loss.backward()
    guidance_train()
    ms_gtrain = timed(guidance_train, 1)
    gtrain_value = world * GH * GW / (ms_gtrain * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- HBM-bound stages at image-sized batches (rank 0, CUDA events per launch, L2 flushed between launches) ---------
    hbm_stages = hbm_stage_times(ops, dev, pk["hbm_gbs"])

    # ---- roofline of the dominant kernel (CUDA events around each launch, inside the timed region) -----
    pts = {"coarse": N_RAND * 64, "fine": N_RAND * 128}
    per_kernel = {}
    tot_kernel_ms = sum(sum(v) for v in ktimes.values())
    for name, vals in ktimes.items():
        per_kernel[name] = {"launches_per_step": len(vals) / a.steps, "ms_per_step": sum(vals) / a.steps,
                            "share_of_step": sum(vals) / a.steps / ms_eager}
    npts = pts["coarse"] + pts["fine"]
    flops = {"mvip_mlp_forward": FLOP_FWD * npts, "dgrad_chain_kernel": FLOP_DGRAD * npts, "wgrad_kernel": FLOP_WGRAD * npts,
             "backward_fused_kernel": (FLOP_DGRAD + FLOP_WGRAD) * npts}
    # algorithmic HBM bytes per point of the training kernels: forward writes the activation stash,
    # the dgrad chain reads the ReLU masks and writes the dZ stash, wgrad reads both stashes (bf16 chunk images only)
    hbm_bytes = {"mvip_mlp_forward": (HBM_FWD_TRAIN, "write"), "dgrad_chain_kernel": (HBM_DGRAD, "write"),
                 "wgrad_kernel": (HBM_WGRAD, "read"), "head_grads_kernel": (HBM_HEADS, "read"),
                 # fused backward: the dZ stash is written once and re-read by the concurrent wgrad CTAs out of L2
                 "backward_fused_kernel": (HBM_DGRAD + 40 * 128, "read+write")}
    top = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"]) if per_kernel else None
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    for k in flops:
        if k in per_kernel:
            per_kernel[k]["tflops"] = flops[k] / (per_kernel[k]["ms_per_step"] * 1e-3) / 1e12
            per_kernel[k]["frac_of_tensor_peak"] = per_kernel[k]["tflops"] / peak_tf
    for k, (b, kind) in hbm_bytes.items():
        if k in per_kernel:
            per_kernel[k]["hbm_gbs"] = b * npts / (per_kernel[k]["ms_per_step"] * 1e-3) / 1e9
            per_kernel[k]["frac_of_hbm_peak"] = per_kernel[k]["hbm_gbs"] / pk["hbm_gbs"]
            per_kernel[k]["hbm_direction"] = kind
    roofline = None
    if top in per_kernel and (top in flops or top in hbm_bytes):
        # every training kernel moves ~5-10 KB per point at < 130 FLOP/B (machine balance ~210): HBM is the binding roof;
        # the tensor-pipe fraction is kept beside it (and is the roof of the render-only forward, see "render")
        e = per_kernel[top]
        if e.get("frac_of_hbm_peak", 0.0) >= e.get("frac_of_tensor_peak", 0.0):
            roofline = {"kernel": top, "bound": "hbm", "achieved": e["hbm_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": e["frac_of_hbm_peak"],
                        "traffic": NCU_TRAFFIC[top] * N_RAND / 4096 if top in NCU_TRAFFIC else None,   # captured at 4096 rays
                        "algorithmic_bytes_per_launch": hbm_bytes[top][0] * npts / 2,
                        "peak_source": "%s hbm_gbs (STREAM-style copy; this kernel's traffic is %s-only, for which "
                                       "the same box measures ~3.9 TB/s write / ~5.9 TB/s read, scripts/hbm_write_bw.py)"
                                       % (pk["_source"], e["hbm_direction"]),
                        "tensor_frac": e.get("frac_of_tensor_peak"), "share_of_step": e["share_of_step"]}
        else:
            roofline = {"kernel": top, "bound": "tensor", "achieved": e["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": e["frac_of_tensor_peak"], "traffic": None,
                        "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % pk["_source"],
                        "share_of_step": e["share_of_step"]}

    # ---- CPU baseline on the host cores (bounded sample) ---------------------------------------------
    cpu_base = None
    if world == 1 and not a.no_cpu_baseline:
        sample, csteps = 512, 4
        cval, _ = cpu_train_rays_per_s(sample, csteps, 1)
        cpu_base = {"value": cval, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": "%d rays x %d steps of the same train step, PyTorch-CPU port of the reference path" % (sample, csteps)}

    bytes_in = (h_o.numel() + h_d.numel() + h_t.numel()) * 4
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": ("cfg4" if N_RAND * world == 65536 else "cfg2") + ": one training step, N_rand=%d rays per GPU (global %d), N_samples=64, N_importance=64, "
                               "coarse+fine 8x256 NeRF (random init), lindisp, white_bkgd, perturb=1, raw_noise_std=1, "
                               "loss=mse(rgb)+mse(rgb0), grad allreduce, fused Adam" % (N_RAND, N_RAND * world),
                   "parallelism": "rays sharded, dp%d" % world,
                   "launch": graph_note,
                   "kernel_times": "CUDA events around every library call in a second timed region of the same %d steps launched "
                                   "eagerly (%.3f ms/step); gpu_launches counts that region" % (a.steps, ms_eager),
                   "l2": "no explicit flush: each step streams ~8 GB of activation stash per GPU (>> 126 MB L2)"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
        "kernels": per_kernel,
        "render": {"value": render_value, "unit": "rays/s", "ms_per_image": ms_render,
                   "workload": "cfg3: 1008x756 image (762,048 rays), render kwargs, rays sharded over %d GPU(s), one gather" % world,
                   "tflops": n_img * 192 * FLOP_FWD / (ms_render * 1e-3) / 1e12,
                   "bound": "tensor", "peak": peak_tf * world,
                   "frac_of_peak": n_img * 192 * FLOP_FWD / (ms_render * 1e-3) / 1e12 / (peak_tf * world)},
        "guidance": {"value": guid_value, "unit": "rays/s", "ms_per_batch": ms_guid,
                     "workload": "cfg5: %d views of %dx%d (rgb + disp + acc + depth), image rows sharded over %d GPU(s), one gather, "
                                 "normal maps (k=31) of all views on rank 0" % (GV, GH, GW, world)},
        "guidance_train": {"value": gtrain_value, "unit": "rays/s", "ms_per_view": ms_gtrain,
                           "workload": "one 512x512 view per GPU, forward + image loss + deferred back-propagation "
                                       "(re-render in 8192-ray chunks), parameter gradients of both networks"},
        "hbm_stages": hbm_stages,
        "train_tflops": world * N_RAND * 192 * (FLOP_FWD + FLOP_DGRAD + FLOP_WGRAD) / (ms_step * 1e-3) / 1e12,
    }
    line["train_frac_of_peak"] = line["train_tflops"] / (peak_tf * world)
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rays-per-gpu", type=int, default=N_RAND,
                    help="N_rand per GPU: 4096 = BASELINE cfg 2 (default, the configuration `metric` is quoted on); 8192 with "
                         "--gpus 8 = cfg 4 (65,536 rays per step)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying CUDA graphs")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
