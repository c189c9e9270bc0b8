#!/usr/bin/env python
"""bench.py -- rays/sec of the Color-NeuS volume-rendering hot path (BASELINE.json metric, config C2).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm: the UNMODIFIED reference on the host cores

One "step" = one pass of the hot path over one batch of synthetic input = rendering every ray of N synthetic 800x800
cameras (640 000 rays per GPU) through Color_NeuS.forward's pipeline: 64 coarse + 64 importance samples, SDF 8x256 +
colour 4x256 + relight 4x256, full return dict produced per chunk.  The flattened ray list (camera-major, y*W+x inside a
camera) is ray-sharded over the ranks with `color_neus_b200.parallel.render_sharded`: contiguous index ranges, no
data-path collective, rgb + depth all-gathered inside the timed region so every rank ends up with every image
(weak scaling: N cameras on N GPUs).  `extra` carries the other BASELINE configs (C3 training step, C4 1080p 128+128,
C5 512^3 extraction + vertex colour), the strong-scaling render of ONE image over the N ranks and the N-rank training
step with its single flat all-reduce.
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

H = W = 800
N_SAMPLES, N_IMPORTANCE = 64, 64
# algorithmic MACs per point from the layer shapes (FLOP = 2*MAC, element-wise work excluded)
MAC_SDF_FULL, MAC_SDF_ONLY, MAC_GRAD, MAC_COLOR, MAC_RELIGHT = 524544, 459008, 459008, 264448, 206592
MAC_SHADE = MAC_SDF_FULL + MAC_GRAD + MAC_COLOR + MAC_RELIGHT


def flop_per_ray(n_s, n_i, train=False):
    """forward = n_so SDF-only points + S fully shaded points; training ~ 3x the shaded part."""
    return 2 * ((n_s + 3 * n_i // 4) * MAC_SDF_ONLY + (3 if train else 1) * (n_s + n_i) * MAC_SHADE)


FLOP_PER_RAY_SHADE = 2 * (N_SAMPLES + N_IMPORTANCE) * MAC_SHADE
FLOP_PER_RAY = flop_per_ray(N_SAMPLES, N_IMPORTANCE)  # 475.2 MFLOP


def renderer_cfg(n_samples=N_SAMPLES, n_importance=N_IMPORTANCE):
    import __graft_entry__ as g
    return g._Cfg(dict(
        TYPE="Color_NeuS", N_SAMPLES=n_samples, N_IMPORTANCE=n_importance, UP_SAMPLE_STEPS=4, PERTURB=1.0,
        SDF=dict(D_IN=3, D_OUT=257, D_HIDDEN=256, N_LAYERS=8, SKIP_IN=[4], MULTIRES=6, BIAS=0.5, SCALE=3.0,
                 GEOMETRIC_INIT=True, WEIGHT_NORM=True, INSIDE_OUTSIDE=False),
        COLOR=dict(D_FEATURE=256, MODE="no_view_dir", D_IN=6, D_OUT=3, D_HIDDEN=256, N_LAYERS=4, WEIGHT_NORM=True,
                   MULTIRES_VIEW=0, SQUEEZE_OUT=True),
        RELIGHT=dict(D_IN=6, D_OUT=3, D_HIDDEN=256, N_LAYERS=4, Y_IN_LAYER=3, MULTIRES_VIEW=4, INCLUDE_GRAD=True,
                     INV_SIGMOID=True),
        DEVIATION=dict(INIT_VAL=0.3)))


def plain(cfg):
    return {k: (plain(v) if isinstance(v, dict) else v) for k, v in cfg.items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    return 1400.0, "B200_PROFILING.md fallback sustained 1.4 PFLOP/s (of fallback)"


def traffic_per_launch(chunk_rays):
    """dram__bytes_read+write of the dominant kernel per launch.  NOT measured in this run (ncu cannot run inside the
    bench): the per-ray figure of the committed `ncu --set full` capture of the same kernel (profiles/traffic.json names
    the capture) scaled to the rays of one launch -- per-tile traffic is constant, so it scales with the launch size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None, None
    d = json.load(open(p))
    return d["dram_bytes_per_ray"] * chunk_rays, f"scaled constant: {d['dram_bytes_per_ray']:.0f} B/ray x {chunk_rays} rays, from {d.get('source', 'profiles/traffic.json')}"


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores
# ---------------------------------------------------------------------------------------------------------------------
class CpuArm:
    """A bounded sample of the workload (centre rows of the 800x800 camera) through the UNMODIFIED reference
    `Color_NeuS.forward` (lib/models/renderers/NeuS.py:294-408 + Color_NeuS.py:24-138), imported from /root/reference or
    its verbatim staged copy oracle/_ref (kind "reference"), autograd on exactly like `validate_image` runs it
    (NeuS_Trainer.py:238-245).  Falls back to the oracle port (kind "port") only where neither tree exists.  All host
    threads, pinned explicitly so that torchrun's OMP_NUM_THREADS=1 cannot change the arm."""

    def __init__(self, state_dict, n_rays, theta, cfg=None):
        import torch
        from color_neus_b200.rays import synthetic_camera_rays
        from oracle import ref_import as R
        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        cfg = plain(cfg if cfg is not None else renderer_cfg())
        ro, rd, near, far = synthetic_camera_rays(H, W, theta_deg=theta, device="cpu")
        s = (H // 2) * W + max(0, (W - n_rays) // 2)
        self.first_ray = s
        self.ro, self.rd = ro[s:s + n_rays].contiguous(), rd[s:s + n_rays].contiguous()
        self.near, self.far = near[s:s + n_rays].contiguous(), far[s:s + n_rays].contiguous()
        self.n_rays = self.ro.shape[0]
        if R.reference_available():
            ns = R.load_reference()
            self.kind = "reference"
            self.what = f"unmodified reference Color_NeuS.forward ({R.reference_kind()} tree), autograd on"
            torch.manual_seed(1)
            self.ren = ns.Color_NeuS(R.CfgDict(cfg))
            if state_dict is not None:
                self.ren.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()}, strict=True)
            self.ren.eval()
            self._run = lambda: self.ren(self.ro, self.rd, self.near, self.far)
        else:
            from oracle import neus_oracle as O
            self.kind = "port"
            self.what = "oracle/neus_oracle.py (restatement; no reference tree on this machine)"
            P = O.to_torch({k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in state_dict.items()}) \
                if state_dict is not None else O.to_torch(O.make_params(cfg, seed=1))

            def run():
                t_rand = torch.rand([self.n_rays, 1])
                with torch.no_grad():
                    return O.render_forward(P, cfg, self.ro, self.rd, self.near, self.far, t_rand=t_rand)
            self._run = run
        self.last = None

    def run_once(self, seed=None):
        if seed is not None:
            self.torch.manual_seed(seed)   # the one CPU draw of the jitter (NeuS.py:325)
        t0 = time.perf_counter()
        out = self._run()
        dt = time.perf_counter() - t0
        self.last = {k: out[k].detach() for k in ("color_fine", "depth", "weight_sum")}
        return dt

    def describe(self):
        return (f"{self.n_rays} centre-row rays (index {self.first_ray}..) of the 800x800 camera per step, 64+64 samples, "
                f"{self.what}, torch CPU fp32 on {self.cores} threads")


def run_reference(args, rank, world):
    """--impl reference: rank 0 only, every step a bounded sample of the workload."""
    if rank != 0:
        return
    import torch
    t0 = time.perf_counter()
    torch.manual_seed(1)
    arm = CpuArm(None, args.ref_rays, 30.0)   # the reference's own constructor under TRAIN.MANUAL_SEED = 1
    for _ in range(args.warmup):
        arm.run_once()
    total = sum(arm.run_once() for _ in range(args.steps))
    value = args.steps * arm.n_rays / total
    cfg = workload_config(args, 1)
    cfg["rays_per_step"] = arm.n_rays
    cfg["sample"] = arm.describe()
    line = {
        "impl": "reference", "metric": "rays/sec (64+64 samples, 256-wide MLP)", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": arm.cores, "kind": arm.kind, "sample": arm.describe()},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "C2: synthetic 800x800 pinhole cameras (640000 rays/GPU/step), Color_NeuS forward, 64+64 "
                        "hierarchical samples, SDF 8x256 + colour 4x256 + relight 4x256, full return dict per chunk",
            "rays_per_step_per_gpu": H * W, "chunk_rays": args.chunk,
            "parallelism": f"ray-sharded x{world} (parallel.render_sharded: contiguous y*W+x ranges of the camera-major ray "
                           "list, all-gather of rgb + depth inside the timed region)",
            "l2": "per-step working set (~3 GB of per-sample outputs) >> 126 MB L2; plus a 256 MiB flush write "
                  "between timed steps"}


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import color_neus_b200 as cn
    from color_neus_b200 import _lib as L
    from color_neus_b200 import parallel as par
    from color_neus_b200.rays import synthetic_camera_rays

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    torch.manual_seed(1)                       # TRAIN.MANUAL_SEED; random-init weights of the named architecture
    ren = cn.Color_NeuS(renderer_cfg()).to(dev).eval()
    thetas = [30.0 + 10.0 * r for r in range(world)]
    cams = [synthetic_camera_rays(H, W, theta_deg=t, device=dev) for t in thetas]
    ro, rd, near, far = (torch.cat([c[i] for c in cams]).contiguous() for i in range(4))   # camera-major ray list
    n_rays = H * W                            # per GPU
    n_total = ro.shape[0]
    g_cpu = torch.Generator().manual_seed(7)
    t_rand = torch.rand([n_total, 1], generator=g_cpu).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    b0, e0 = par.shard_range(n_total, rank, world)

    def render_fn(o, d, n_, f_, t_rand):
        return ren._forward_impl(o, d, n_, f_, t_rand=t_rand)

    def step_resident():
        out, _ = par.render_sharded(render_fn, ro, rd, near, far, gather=("color_fine", "depth"), chunk=args.chunk,
                                    keep=("color_fine", "depth"), per_ray_kw={"t_rand": t_rand})
        return out

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        lib.cneus_profile_enable(1)
        ms0, n0 = C.c_double(), C.c_int64()
        lib.cneus_profile_read(0, C.byref(ms0), C.byref(n0))
        lib.cneus_profile_read(1, C.byref(ms0), C.byref(n0))
        launches0 = lib.cneus_launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for a, b in evs:
            flush.zero_()                      # evict L2 between timed steps (not inside the per-step events)
            a.record()
            image = step_resident()
            b.record()
        barrier()
        launches = lib.cneus_launch_count() - launches0
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        lib.cneus_profile_enable(0)
        ms_full, n_full, ms_only, n_only = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
        lib.cneus_profile_read(1, C.byref(ms_full), C.byref(n_full))
        lib.cneus_profile_read(0, C.byref(ms_only), C.byref(n_only))
        clk = clocks.stop()

        # ---- end to end through the public API: host buffers in, host results out, every step (this rank's slice)
        ro_h, rd_h = ro[b0:e0].cpu().pin_memory(), rd[b0:e0].cpu().pin_memory()
        near_h, far_h = near[b0:e0].cpu().pin_memory(), far[b0:e0].cpu().pin_memory()
        n_loc = e0 - b0
        color_h = torch.empty(n_loc, 3).pin_memory()
        depth_h = torch.empty(n_loc).pin_memory()
        chunks = [(s, min(s + args.chunk, n_loc)) for s in range(0, n_loc, args.chunk)]

        def step_e2e():
            for s, e in chunks:
                r = ren(ro_h[s:e].to(dev, non_blocking=True), rd_h[s:e].to(dev, non_blocking=True),
                        near_h[s:e].to(dev, non_blocking=True), far_h[s:e].to(dev, non_blocking=True))
                color_h[s:e].copy_(r["color_fine"], non_blocking=True)
                depth_h[s:e].copy_(r["depth"], non_blocking=True)
            torch.cuda.synchronize()

        e2e_s = 1.0
        if args.e2e_steps > 0:
            step_e2e()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                step_e2e()
            barrier()
            e2e_s = time.perf_counter() - t0

    total_ms, e2e_s = max_over_ranks(total_ms, e2e_s)
    value = n_total * args.steps / (total_ms * 1e-3)
    e2e_value = n_total * args.e2e_steps / e2e_s

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        shade_rays = (e0 - b0) * args.steps
        achieved = (shade_rays * FLOP_PER_RAY_SHADE / max(ms_full.value, 1e-9) / 1e9) if n_full.value else None  # TFLOP/s
        traffic, traffic_src = traffic_per_launch(args.chunk)
        line = {
            "metric": "rays/sec (64+64 samples, 256-wide MLP)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "rays/s",
                    "h2d_bytes_per_step": n_rays * (3 + 3 + 1 + 1 + 1) * 4, "d2h_bytes_per_step": n_rays * 4 * 4,
                    "bytes_are": "per GPU", "steps": args.e2e_steps,
                    "api": "color_neus_b200.Color_NeuS.forward (pinned host rays in, colour+depth to pinned host out, "
                           "per chunk; every rank its contiguous slice)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor",
                         "kernel": "shade_tc_kernel, render_core launch (SDF + gradient chain + colour + relight on "
                                   "tcgen05 as CTA pairs (cta_group::2), fp16 hi/lo 3-pass = 3 MMAs per algorithmic MAC)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "flop_per_launch": FLOP_PER_RAY_SHADE * args.chunk, "launches": int(n_full.value),
                         "avg_launch_ms": ms_full.value / max(n_full.value, 1),
                         "share_of_step": ms_full.value / total_ms,
                         "sdf_only_launch_ms_total": ms_only.value, "sdf_only_launches": int(n_only.value),
                         "whole_step_tflops": n_total * args.steps * FLOP_PER_RAY / (total_ms * 1e-3) / 1e12,
                         "counters_are": "rank 0's launches"},
        }
        if world == 1 and not args.no_cpu_baseline:
            # ---- the reference on this box's host cores, same weights, same camera; and parity of the benched image
            arm = CpuArm(ren.state_dict(), args.ref_rays, thetas[0])
            arm.run_once(seed=7)
            times = sorted(arm.run_once(seed=7) for _ in range(2))
            rps = arm.n_rays / times[len(times) // 2]
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": arm.cores, "kind": arm.kind,
                                    "sample": "median of 2 runs (after 1 warm-up) over " + arm.describe() + ", same weights"}
            s = arm.first_ray
            torch.manual_seed(7)              # same CPU jitter draw as the arm's last run
            with torch.no_grad():
                got = ren(ro[s:s + arm.n_rays], rd[s:s + arm.n_rays], near[s:s + arm.n_rays], far[s:s + arm.n_rays])

            def rel(k):
                a, b = got[k].float().cpu().reshape(-1), arm.last[k].float().reshape(-1)
                return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
            line["parity"] = {"against": arm.kind, "rays": arm.n_rays, "first_ray_index": s, "metric": "max|err| / max|ref|",
                              "max_rel_err_rgb": rel("color_fine"), "depth": rel("depth"), "weight_sum": rel("weight_sum"),
                              "tolerance_rgb": 1e-4}
    # ---- the other BASELINE configs and the multi-GPU paths (every rank takes part; rank 0 reports)
    extra = {}
    if args.extras != "none":
        ctx = dict(torch=torch, dist=dist, cn=cn, par=par, L=L, lib=lib, dev=dev, rank=rank, world=world, barrier=barrier,
                   max_over_ranks=max_over_ranks, args=args, ren=ren, image=image if world > 1 else None)
        for name, fn in (("train_step", extra_train_step), ("strong", extra_strong), ("c4", extra_c4), ("c5", extra_c5),
                         ("reference_on_b200", extra_reference_on_gpu)):
            if args.extras not in ("all", name) and name not in args.extras.split(","):
                continue
            try:
                r = fn(ctx)
                if r is not None:
                    extra[r.pop("key", name)] = r
            except Exception as e:   # an extra must not take the contract line down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
                if world > 1:
                    raise
    if rank == 0:
        line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_reference_on_gpu(c):
    """Context, not a baseline arm: the UNMODIFIED reference renderer (oracle/_ref or /root/reference) moved to the same B200
    with `.cuda()` -- fp32, TF32 left off as the reference leaves it, autograd on as validate_image runs it -- on rays of the
    same camera, in the reference's own EVAL_RAY_SIZE = 1024 chunks and in 8192-ray chunks.  Says how much of the speed-up
    is "a GPU at all" (ATen / cuBLAS SGEMM kernels, ~45 activation round trips through HBM per point) and how much is this
    repo's fused kernels."""
    torch, dev, rank, world, ren = c["torch"], c["dev"], c["rank"], c["world"], c["ren"]
    if world > 1 or rank != 0:
        return None
    from oracle import ref_import as R
    if not R.reference_available():
        return {"unavailable": "no reference tree on this machine"}
    from color_neus_b200.rays import synthetic_camera_rays
    ns = R.load_reference()
    ref = ns.Color_NeuS(R.CfgDict(plain(renderer_cfg())))
    ref.load_state_dict({k: v.detach().cpu() for k, v in ren.state_dict().items()}, strict=True)
    ref = ref.to(dev).eval()
    ro, rd, near, far = synthetic_camera_rays(H, W, theta_deg=30.0, device=dev)
    s0 = (H // 2 - 8) * W
    out = {"what": "unmodified reference Color_NeuS.forward on cuda:0 (ATen kernels, fp32, autograd on), centre rows of the bench camera"}
    for chunk, n_chunks in ((1024, 8), (8192, 2)):
        def run():
            for i in range(n_chunks):
                a, b = s0 + i * chunk, s0 + (i + 1) * chunk
                r = ref(ro[a:b], rd[a:b], near[a:b], far[a:b])
                r["color_fine"].detach()
        run()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); run(); ev1.record()
        torch.cuda.synchronize()
        out[f"rays_per_s_chunk_{chunk}"] = chunk * n_chunks / (ev0.elapsed_time(ev1) * 1e-3)
    torch.manual_seed(7)
    got_ref = ref(ro[s0:s0 + 1024], rd[s0:s0 + 1024], near[s0:s0 + 1024], far[s0:s0 + 1024])
    torch.manual_seed(7)
    with torch.no_grad():
        got = ren(ro[s0:s0 + 1024], rd[s0:s0 + 1024], near[s0:s0 + 1024], far[s0:s0 + 1024])
    a, b = got["color_fine"].float(), got_ref["color_fine"].detach().float()
    out["max_rel_err_rgb_vs_this_repo"] = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    del ref
    torch.cuda.empty_cache()
    return out


def extra_train_step(c):
    """BASELINE C3: one training step = N_RAYS=1024 rays per GPU of a 768x576 camera, 64+128 samples, Color_NeuS forward +
    NeuS_Trainer.compute_loss + backward + per-tensor clip + Adam (train.py:63-77).  At N > 1 the union batch of 1024 N
    rays is ray-sharded: union-batch loss from all-reduced partial sums (3 scalars), backward, ONE all-reduce of the flat
    gradient buffer (+ the loss share), fused clip + Adam identically on every rank."""
    torch, dist, cn, par, dev, rank, world = c["torch"], c["dist"], c["cn"], c["par"], c["dev"], c["rank"], c["world"]
    from color_neus_b200 import train_ops as TR
    from color_neus_b200.rays import synthetic_camera_rays
    n_s, n_i, n_per = 64, 128, 1024
    n_union = n_per * world
    torch.manual_seed(1)
    ren = cn.Color_NeuS(renderer_cfg(n_s, n_i)).to(dev).train()
    ro, rd, near, far = synthetic_camera_rays(768, 576, device=dev)
    gen = torch.Generator().manual_seed(3)
    idx = torch.randint(0, ro.shape[0], (n_union,), generator=gen).to(dev)
    gt = torch.rand(n_union, 3, generator=gen).to(dev)
    b, e = par.shard_range(n_union, rank, world)
    sl = idx[b:e]
    ro, rd, near, far, gt = ro[sl].contiguous(), rd[sl].contiguous(), near[sl].contiguous(), far[sl].contiguous(), gt[b:e]
    opt = TR.FusedClipAdam(ren.parameters(), lr=5e-4, betas=(0.9, 0.99))
    buf = par.FlatGradBuffer(ren.parameters(), n_extra=1)
    ev_ar = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step():
        buf.zero()
        r = ren(ro, rd, near, far)
        mask = (r["weight_sum"].detach().squeeze(-1) > 0.5).float()
        sums = par.loss_partial_sums(r, mask)
        if world > 1:
            dist.all_reduce(sums)
        loss = par.union_batch_loss(r, gt, mask, n_union, sums)
        loss.backward()
        buf.extra.copy_(loss.detach().reshape(1))
        ev_ar[0].record()
        buf.all_reduce()
        ev_ar[1].record()
        TR.clip_gradient(opt, 1.0, 2)
        opt.step()

    for _ in range(3):
        step()
    c["barrier"]()
    torch.cuda.reset_peak_memory_stats(dev)
    steps = 5
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ar_ms = []
    for a, b_ in ev:
        a.record(); step(); b_.record()
        torch.cuda.synchronize()
        ar_ms.append(ev_ar[0].elapsed_time(ev_ar[1]))
    c["barrier"]()
    ms = sorted(a.elapsed_time(b_) for a, b_ in ev)[steps // 2]
    ms, ar = c["max_over_ranks"](ms, sorted(ar_ms)[steps // 2])
    peak, _ = measured_peaks()
    tflops = n_union * flop_per_ray(n_s, n_i, train=True) / ms / 1e9
    out = {"key": "train_step" if world == 1 else "train_step_ddp",
           "workload": f"C3: {n_per} rays/GPU x {world} GPU(s), 768x576 camera, {n_s}+{n_i} samples, Color_NeuS fwd + loss + bwd + "
                       "clip + Adam (fused loss / clip / Adam kernels, persistent flat gradient buffer)",
           "ms_per_step": ms, "rays_per_s": n_union / ms * 1e3, "algorithmic_tflops": tflops, "frac_of_peak": tflops / (peak * world),
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30, "timing": "median of 5 steps after 3 warm-up, CUDA events, max over ranks"}
    if world > 1:
        out["allreduce_ms"] = ar
        out["collectives_per_step"] = "all_reduce(3 loss partial sums) + all_reduce(flat gradients + loss share: " \
                                      f"{buf.flat.numel()} floats)"
    return out if rank == 0 else None


def extra_strong(c):
    """Strong scaling: ONE 800x800 image ray-sharded over the N ranks (all-gather of rgb + depth in the timed region)."""
    torch, par, dev, rank, world, ren, args = c["torch"], c["par"], c["dev"], c["rank"], c["world"], c["ren"], c["args"]
    if world == 1:
        return None
    from color_neus_b200.rays import synthetic_camera_rays
    ro, rd, near, far = synthetic_camera_rays(H, W, theta_deg=30.0, device=dev)
    t_rand = torch.rand([H * W, 1], generator=torch.Generator().manual_seed(7)).to(dev)
    chunk = min(args.chunk, -(-H * W // world))

    def once():
        out, _ = par.render_sharded(lambda o, d, n_, f_, t_rand: ren._forward_impl(o, d, n_, f_, t_rand=t_rand), ro, rd, near, far,
                                    gather=("color_fine", "depth"), chunk=chunk, keep=("color_fine", "depth"),
                                    per_ray_kw={"t_rand": t_rand})
        return out
    with torch.no_grad():
        once()
        c["barrier"]()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in ev:
            a.record(); out = once(); b.record()
        c["barrier"]()
    ms, = c["max_over_ranks"](sum(a.elapsed_time(b) for a, b in ev) / 3)
    # the sharded image equals the weak-scaling run's camera 0 (same rays, same jitter): sharding invariance on the GPUs
    same = None
    if c["image"] is not None:
        same = bool(torch.equal(out["color_fine"], c["image"]["color_fine"][:H * W]) and torch.equal(out["depth"], c["image"]["depth"][:H * W]))
    return {"workload": f"one 800x800 image (640000 rays) over {world} ranks, contiguous y*W+x ranges, all-gather of rgb + depth",
            "ms_per_image": ms, "rays_per_s": H * W / ms * 1e3, "bit_identical_to_camera0_of_the_weak_run": same} if rank == 0 else None


def extra_c4(c):
    """BASELINE C4: 1920x1080 (2 073 600 rays), 128+128 samples, relight branch on, rays sharded over the N ranks."""
    torch, cn, par, dev, rank, world, args = c["torch"], c["cn"], c["par"], c["dev"], c["rank"], c["world"], c["args"]
    from color_neus_b200.rays import synthetic_camera_rays
    hh, ww, n_s, n_i = 1080, 1920, 128, 128
    torch.manual_seed(1)
    ren = cn.Color_NeuS(renderer_cfg(n_s, n_i)).to(dev).eval()
    ro, rd, near, far = synthetic_camera_rays(hh, ww, theta_deg=30.0, device=dev)
    t_rand = torch.rand([hh * ww, 1], generator=torch.Generator().manual_seed(7)).to(dev)
    chunk = min(args.chunk // 2, 16384)

    def once(n):
        out, _ = par.render_sharded(lambda o, d, n_, f_, t_rand: ren._forward_impl(o, d, n_, f_, t_rand=t_rand), ro[:n], rd[:n], near[:n],
                                    far[:n], gather=("color_fine", "depth"), chunk=chunk, keep=("color_fine", "depth"),
                                    per_ray_kw={"t_rand": t_rand[:n]})
        return out
    with torch.no_grad():
        once(chunk * world * 2)      # warm-up: packing, workspace, NCCL buffers
        c["barrier"]()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); once(hh * ww); b.record()
        c["barrier"]()
    ms, = c["max_over_ranks"](a.elapsed_time(b))
    peak, _ = measured_peaks()
    tf = hh * ww * flop_per_ray(n_s, n_i) / ms / 1e9
    return {"workload": f"C4: 1920x1080 ({hh * ww} rays), 128+128 samples, Color_NeuS (relight on), ray-sharded over {world} rank(s), "
                        "all-gather of rgb + depth", "ms_per_frame": ms, "rays_per_s": hh * ww / ms * 1e3, "algorithmic_tflops": tf,
            "frac_of_peak": tf / (peak * world), "timing": "one full frame after a warm-up, CUDA events, max over ranks"} if rank == 0 else None


def extra_c5(c):
    """BASELINE C5: 512^3 SDF grid (slab-sharded, one all-gather) -> device marching cubes -> per-vertex colour
    (vertex-sharded, one all-gather)."""
    torch, par, dev, rank, world, ren = c["torch"], c["par"], c["dev"], c["rank"], c["world"], c["ren"]
    from color_neus_b200.marching_cubes import marching_cubes_device
    res = 512
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    with torch.no_grad():
        marching_cubes_device(ren.extract_fields(bmin, bmax, 32).reshape(32, 32, 32), 0.0)   # warm-up (tables, workspace)
        if world > 1:
            par.extract_fields_sharded(lambda b_, e_: ren.extract_fields(bmin, bmax, 32, b_, e_), 32)
        c["barrier"]()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        u = par.extract_fields_sharded(lambda b_, e_: ren.extract_fields(bmin, bmax, res, b_, e_), res)
        ev[1].record()
        v, f = marching_cubes_device(u.reshape(res, res, res), 0.0)
        ev[2].record()
        vw = (v / (res - 1.0) * (bmax - bmin).to(dev).double() + bmin.to(dev).double()).float().contiguous()

        def color_fn(pts):
            rgb = torch.empty(pts.shape[0], 3, device=dev)
            h = ren.handle()
            ws, wsb = h.workspace(n_points=pts.shape[0])
            L = c["L"]
            L.check(c["lib"].cneus_vertex_color(h.dref(), h.packed(), L.ptr(pts.contiguous()), pts.shape[0], L.ptr(rgb), ws, wsb,
                                                L.stream_ptr()), "cneus_vertex_color")
            return rgb
        col = par.extract_color_sharded(color_fn, vw)
        ev[3].record()
        c["barrier"]()
    grid_ms, mc_ms, col_ms = c["max_over_ranks"](ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]))
    n = res ** 3
    return {"workload": f"C5: {res}^3 SDF grid in slabs over {world} rank(s) + all-gather, marching cubes on the device, per-vertex colour "
                        f"over {world} rank(s) + all-gather", "grid_ms": grid_ms, "grid_points_per_s": n / grid_ms * 1e3,
            "grid_algorithmic_tflops": n * 2 * MAC_SDF_ONLY / grid_ms / 1e9, "marching_cubes_ms": mc_ms, "vertices": int(v.shape[0]),
            "triangles": int(f.shape[0]), "vertex_color_ms": col_ms, "colors_finite": bool(torch.isfinite(col).all()),
            "timing": "single pass after a 32^3 warm-up, CUDA events, max over ranks"} if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=32768, help="rays per renderer call")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-rays", type=int, default=512, help="rays per CPU-arm sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default="all", help="all | none | comma list of train_step,strong,c4,c5,reference_on_b200")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
