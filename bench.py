#!/usr/bin/env python
"""bench.py -- rays/sec of the Color-NeuS volume-rendering hot path (BASELINE.json metric, config C2).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port of the reference)

One "step" = one pass of the hot path over one batch of synthetic input = rendering every ray of one synthetic
800x800 camera (640 000 rays) through Color_NeuS.forward's pipeline: 64 coarse + 64 importance samples,
SDF 8x256 + colour 4x256 + relight 4x256, full return dict produced per chunk.  Rays are independent, so at N>1
every rank renders its own camera (weak scaling, no data-path collective).
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
FLOP_PER_RAY_SHADE = 2 * (N_SAMPLES + N_IMPORTANCE) * (MAC_SDF_FULL + MAC_GRAD + MAC_COLOR + MAC_RELIGHT)
FLOP_PER_RAY_SAMPLING = 2 * (N_SAMPLES + 3 * N_IMPORTANCE // 4) * MAC_SDF_ONLY
FLOP_PER_RAY = FLOP_PER_RAY_SHADE + FLOP_PER_RAY_SAMPLING  # 475.2 MFLOP


def renderer_cfg():
    import __graft_entry__ as g
    return g._Cfg(dict(
        TYPE="Color_NeuS", N_SAMPLES=N_SAMPLES, N_IMPORTANCE=N_IMPORTANCE, UP_SAMPLE_STEPS=4, PERTURB=1.0,
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
    """dram__bytes_read+write of the dominant kernel per launch, from the committed ncu --set full capture
    (profiles/traffic.json; per-tile traffic is constant, so it scales with the rays of a launch)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None
    return json.load(open(p))["dram_bytes_per_ray"] * chunk_rays


class CpuArm:
    """The oracle port of the reference on the host cores: a bounded sample of the same camera (centre rows)."""

    def __init__(self, state_dict_np, n_rays, theta):
        import torch
        from oracle import neus_oracle as O
        self.O, self.torch = O, torch
        self.cfg = plain(renderer_cfg())
        self.P = O.to_torch(state_dict_np)
        c2w = O.pose_spherical(theta, -30.0, 2.8)
        ro, rd = O.get_rays_at(c2w, torch.tensor([1.2 * W, 1.2 * W]), H, W)
        s = (H // 2) * W + max(0, (W - n_rays) // 2)
        self.ro, self.rd = ro[s:s + n_rays].contiguous(), rd[s:s + n_rays].contiguous()
        self.near, self.far = O.near_far_from_sphere(self.ro, self.rd)
        self.n_rays = self.ro.shape[0]
        self.cores = torch.get_num_threads()
        torch.manual_seed(7)

    def run_once(self):
        t_rand = self.torch.rand([self.n_rays, 1])
        t0 = time.perf_counter()
        with self.torch.no_grad():
            self.O.render_forward(self.P, self.cfg, self.ro, self.rd, self.near, self.far, t_rand=t_rand)
        return time.perf_counter() - t0


def cpu_oracle_rays_per_s(state_dict_np, n_rays, repeats, theta):
    arm = CpuArm(state_dict_np, n_rays, theta)
    arm.run_once()
    times = sorted(arm.run_once() for _ in range(repeats))
    return arm.n_rays / times[len(times) // 2], arm.cores


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the Python reference
    itself cannot travel to the GPU box); rank 0 only, every step a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import neus_oracle as O
    t0 = time.perf_counter()
    arm = CpuArm(O.make_params(plain(renderer_cfg()), seed=1), args.ref_rays, 30.0)
    for _ in range(args.warmup):
        arm.run_once()
    total = sum(arm.run_once() for _ in range(args.steps))
    value = args.steps * arm.n_rays / total
    line = {
        "impl": "reference", "metric": "rays/sec (64+64 samples, 256-wide MLP)", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": arm.cores, "kind": "port",
                         "sample": f"{arm.n_rays} centre-row rays of the 800x800 camera per step, oracle/neus_oracle.py "
                                   f"(torch CPU fp32, {arm.cores} threads)"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "C2: synthetic 800x800 pinhole camera (640000 rays/GPU/step), Color_NeuS forward, 64+64 "
                        "hierarchical samples, SDF 8x256 + colour 4x256 + relight 4x256, full return dict per chunk",
            "rays_per_step_per_gpu": H * W, "chunk_rays": args.chunk, "parallelism": f"ray-sharded x{world}",
            "l2": "per-step working set (~3 GB of per-sample outputs) >> 126 MB L2; plus a 256 MiB flush write "
                  "between timed steps"}


def run_ours(args, rank, world, local_rank):
    import torch
    import __graft_entry__ as g
    g.build()
    import color_neus_b200 as cn
    from color_neus_b200 import _lib as L
    from color_neus_b200.rays import synthetic_camera_rays

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()

    torch.manual_seed(1)                       # TRAIN.MANUAL_SEED; random-init weights of the named architecture
    ren = cn.Color_NeuS(renderer_cfg()).to(dev).eval()
    theta = 30.0 + 10.0 * rank
    ro, rd, near, far = synthetic_camera_rays(H, W, theta_deg=theta, device=dev)
    n_rays = ro.shape[0]
    g_cpu = torch.Generator().manual_seed(7 + rank)
    t_rand_h = torch.rand([n_rays, 1], generator=g_cpu)
    t_rand = t_rand_h.to(dev)
    color = torch.empty(n_rays, 3, device=dev)
    depth = torch.empty(n_rays, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    chunks = [(s, min(s + args.chunk, n_rays)) for s in range(0, n_rays, args.chunk)]

    def step_resident():
        for s, e in chunks:
            r = ren._forward_impl(ro[s:e], rd[s:e], near[s:e], far[s:e], t_rand=t_rand[s:e])
            color[s:e] = r["color_fine"]
            depth[s:e] = r["depth"]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        lib.cneus_profile_enable(1)
        import ctypes as C
        ms0, n0 = C.c_double(), C.c_int64()
        lib.cneus_profile_read(0, C.byref(ms0), C.byref(n0))
        lib.cneus_profile_read(1, C.byref(ms0), C.byref(n0))
        launches0 = lib.cneus_launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for a, b in evs:
            flush.zero_()                      # evict L2 between timed steps (not inside the per-step events)
            a.record()
            step_resident()
            b.record()
        barrier()
        launches = lib.cneus_launch_count() - launches0
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        lib.cneus_profile_enable(0)
        ms_full, n_full, ms_only, n_only = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
        lib.cneus_profile_read(1, C.byref(ms_full), C.byref(n_full))
        lib.cneus_profile_read(0, C.byref(ms_only), C.byref(n_only))
        clk = clocks.stop()

        # ---- end to end through the public API: host buffers in, host results out, every step
        ro_h, rd_h = ro.cpu().pin_memory(), rd.cpu().pin_memory()
        near_h, far_h = near.cpu().pin_memory(), far.cpu().pin_memory()
        color_h = torch.empty(n_rays, 3).pin_memory()
        depth_h = torch.empty(n_rays).pin_memory()

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

    t_max = torch.tensor([total_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t_max[0]), float(t_max[1])
    value = world * n_rays * args.steps / (total_ms * 1e-3)
    e2e_value = world * n_rays * args.e2e_steps / e2e_s

    if rank == 0:
        peak, peak_src = measured_peaks()
        shade_rays = n_rays * args.steps
        achieved = (shade_rays * FLOP_PER_RAY_SHADE / max(ms_full.value, 1e-9) / 1e9) if n_full.value else None  # TFLOP/s
        line = {
            "metric": "rays/sec (64+64 samples, 256-wide MLP)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "rays/s",
                    "h2d_bytes_per_step": n_rays * (3 + 3 + 1 + 1 + 1) * 4, "d2h_bytes_per_step": n_rays * 4 * 4,
                    "steps": args.e2e_steps, "api": "color_neus_b200.Color_NeuS.forward (pinned host rays in, "
                                                     "colour+depth to pinned host out, per chunk)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor",
                         "kernel": "shade_tc_kernel, render_core launch (SDF + gradient chain + colour + relight on "
                                   "tcgen05, fp16 hi/lo 3-pass = 3 MMAs per algorithmic MAC)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic_per_launch(args.chunk),
                         "peak_source": peak_src,
                         "flop_per_launch": FLOP_PER_RAY_SHADE * args.chunk, "launches": int(n_full.value),
                         "avg_launch_ms": ms_full.value / max(n_full.value, 1),
                         "share_of_step": ms_full.value / total_ms,
                         "sdf_only_launch_ms_total": ms_only.value, "sdf_only_launches": int(n_only.value),
                         "whole_step_tflops": world * n_rays * args.steps * FLOP_PER_RAY / (total_ms * 1e-3) / 1e12},
        }
        if world == 1 and not args.no_cpu_baseline:
            sd = {k: v.detach().cpu().numpy() for k, v in ren.state_dict().items()}
            rps, cores = cpu_oracle_rays_per_s(sd, args.ref_rays, 2, theta)
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                                    "sample": f"median of 2 runs (after 1 warm-up) over {args.ref_rays} centre-row rays "
                                              f"of the same camera, same weights; oracle/neus_oracle.py on {cores} torch threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


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
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
