#!/usr/bin/env python
"""Benchmark of the guided denoising loop (BASELINE.json metric: guided img-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], per GPU): DDPM-256 UNet2DModel (random-init, seed 0), batch 8,
regeneration steps of the edit-friendly DDPM inversion (reverse_step, eta = 1, extracted-style noise
maps z_t) with SingleColorAttrFunc(target 0.8, channel 0, loss_scale 50) guidance after every step,
run through the drop-in ``SegDiffEditPipeline.edit_image`` (inversion_method="ddpm", Tskip).
One "step" = one guided denoising step of the whole per-GPU batch: UNet forward (tcgen05 kernels)
+ one fused step kernel.  value = images x steps / second over all GPUs (weak scaling: the batch is
sharded per GPU, no data-path collective; one NCCL all_gather of the final images at the end).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "diffusion-image-editing_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "guided img-steps/s (DDPM-256 UNet, edit-friendly DDPM regeneration step + colour guidance)"
UNIT = "img-steps/s"
TARGET, CHANNEL, LOSS_SCALE, ETA, T_INFER = 0.8, 0, 50.0, 1.0, 50


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sust=float(p["bf16_tflops_sustained"]), src="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region (rank 0's GPU)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, enabled=True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:     # only the reporting rank samples: N polling nvidia-smi processes perturb an N-rank run
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.th.join(timeout=1)

    def summary(self):
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_guided_steps(n_steps, warmup, batch=1, budget_s=120.0):
    """The reference's CPU implementation of the step (oracle port: torch-CPU UNet2DModel restatement
    + reference step math), all host threads.  Returns (img-steps/s, timed steps, seconds/step)."""
    from oracle import step_math as sm
    from oracle.ddim_scheduler import DDIMScheduler
    from oracle.unet2d import DDPM256_CONFIG, UNet2DModel
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    unet = UNet2DModel(**DDPM256_CONFIG).eval()
    sch = DDIMScheduler.from_preset("ddpm", clip_sample=False)
    sch.set_timesteps(T_INFER)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(batch, 3, 256, 256, generator=g)
    ts = [int(t) for t in sch.timesteps]
    guidance_targets = [TARGET if c == CHANNEL else None for c in range(3)]

    def one(i, x):
        t = ts[(len(ts) - 14 + i) % len(ts)]
        with torch.no_grad():
            eps = unet(x, torch.tensor(t))["sample"]
        c = sm.step_coeffs(sch, t)
        z = torch.randn(3, 256, 256, generator=g)
        xp, _ = sm.ddpm_reverse_step(x, eps, c, ETA, z)
        xp, _ = sm.color_guidance_update(xp, eps, c, guidance_targets, [1, 1, 1], LOSS_SCALE)
        return xp

    t0 = time.perf_counter()
    for i in range(max(1, warmup)):
        x = one(i, x)
    per = (time.perf_counter() - t0) / max(1, warmup)
    n = max(1, min(n_steps, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for i in range(n):
        x = one(i, x)
    dt = time.perf_counter() - t0
    return batch * n / dt, n, dt / n


def run_reference(args, rank):
    if rank != 0:
        return
    v, n, per = cpu_guided_steps(args.steps, min(args.warmup, 1), batch=1, budget_s=150.0)
    cores = torch.get_num_threads()
    sample = (f"oracle port of the reference CPU path (torch-CPU fp32 UNet2DModel restatement + reference step "
              f"math), batch 1 per step, {n} timed steps of the {args.steps} requested ({per:.2f} s/step), {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch), "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores,
                                                                 "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {"workload": "BASELINE configs[1]: DDPM-256 UNet2DModel (random-init) colour-guided regeneration step of "
                        "the edit-friendly DDPM inversion (reverse_step eta=1 + SingleColorAttrFunc), per-GPU batch "
                        f"{batch}", "batch_per_gpu": batch, "global_batch": batch * args.gpus, "image": "3x256x256",
            "num_inference_steps": T_INFER, "eta": ETA, "guidance": "SingleColorAttrFunc(target=0.8,color_idx=0,loss_scale=50)",
            "parallelism": f"dp{args.gpus} (batch sharded, final NCCL all_gather)",
            "l2": "per-step working set (UNet activations, GBs) exceeds the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------- native arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--step-kernel-batch", type=int, default=256)
    ap.add_argument("--profile-out", default=None, help="write the per-op CUDA-event profile (JSON) here")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="noise predictor: bf16 operands (headline) or the fp32-accurate split-bf16 mode")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from attr_functions import SingleColorAttrFunc
    from b200edit import _C, ops
    from b200edit.distributed import gather_images
    from models import create_diffusion_model
    from SegDiffEditPipeline import SegDiffEditPipeline

    B, K, W = args.batch, args.steps, args.warmup
    wrapper = create_diffusion_model("ddpm", sample_clipping=False, max_batch=B, seed=0, precision=args.precision)
    sch = wrapper.scheduler
    T = T_INFER if K <= T_INFER else K
    sch.set_timesteps(T)
    pipe = SegDiffEditPipeline(wrapper, None)
    f = SingleColorAttrFunc(target=TARGET, color_idx=CHANNEL, loss_scale=LOSS_SCALE, t1=0, t2=10 ** 9, per_sample=True)
    gen = torch.Generator().manual_seed(1000 + rank)        # per-rank shard of the global batch
    S = 256
    # host (pinned) copies of the inputs: x_T shard and the per-step noise maps of every sample
    xts_host = torch.randn(T + 1, 3, S, S, generator=gen).pin_memory()      # reference layout: xts[Tskip] is the start
    x_host = torch.randn(B, 3, S, S, generator=gen).pin_memory()
    zs_host = torch.randn(K, 3, S, S, generator=gen).pin_memory()           # (C,H,W) per step, broadcast over the batch
    out_host = torch.empty(B, 3, S, S).pin_memory()
    x0_host = torch.empty(K, B, 3, S, S).pin_memory()
    del xts_host

    def run(K_, from_host):
        """K_ guided steps through the public pipeline call.  from_host: H2D of x_T / z inside, D2H of results."""
        if from_host:
            x = x_host.to(dev, non_blocking=True)
            zs = zs_host[:K_].to(dev, non_blocking=True)
        else:
            x, zs = x_dev, zs_dev[:K_]
        # host leg: the x0-prediction history streams to pinned host memory step by step on the pipeline's copy stream
        # (overlapped with the next step's UNet forward, joined before the call returns); the final images follow
        out = pipe.edit_image(xt=x, eta=ETA, zs=zs, attr_func=f, inversion_method="ddpm", Tskip=0, xts=None,
                              prog_bar=False, output_type="tensor", x0_history_out=x0_host[:K_] if from_host else None)
        if from_host:
            out_host.copy_(out.imgs, non_blocking=True)
        return out

    # NOTE edit_image with xts=None keeps xt as the start sample; Tskip only selects the ddpm reverse_step branch.
    x_dev = x_host.to(dev)
    zs_dev = zs_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(K_, from_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _C.launch_count()
        e0.record()
        out = run(K_, from_host)
        if world > 1:
            gather_images(out.imgs, world * B)     # the path's only collective: final result gather (NCCL)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _C.launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], device=dev, dtype=torch.int64)
            dist.all_reduce(lt)
            launches = int(lt.item())
        return ms, launches

    run(W, False)          # warm-up: W untimed steps (plans, function attributes)
    run(min(W, K), True)
    # one untimed pass of the timed shape as well: edit_image returns the K-long x0 history, so the first
    # K-step call grows torch's caching allocator (cudaMalloc inside the timed region otherwise)
    run(K, False)
    warm = run(K, True)
    if world > 1:
        gather_images(warm.imgs, world * B)    # untimed first collective: NCCL sets its channels up lazily
    del warm
    with ClockSampler(local_rank, enabled=(rank == 0)) as clk:
        ms_dev, launches = timed(K, False)
        ms_e2e, _ = timed(K, True)
    value = world * B * K / (ms_dev * 1e-3)
    e2e = world * B * K / (ms_e2e * 1e-3)

    if rank == 0:
        pk = peaks()
        # ---- roofline of the dominant kernel (tensor bound): the tcgen05 implicit-GEMM convolution.  `roofline` is
        # the kernel variant with the largest share of the step (halo, 2 M tiles per CTA: the 256x256 layers);
        # `roofline_unet_convs` is every conv / linear / attention GEMM launch of the forward together.
        prof = wrapper.unet.profile(x_dev, int(sch.timesteps[-1]))
        conv = [p for p in prof if p["kind"] == "conv_igemm"]
        top = [p for p in conv if "halo2" in p["desc"] and "bn128" in p["desc"]] or conv
        tot_ms = sum(p["ms"] for p in prof)
        by_kind = {}
        for p in prof:
            d = by_kind.setdefault(p["kind"], {"ms": 0.0, "launch_groups": 0})
            d["ms"] += p["ms"]
            d["launch_groups"] += 1
        try:
            with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
                ncu = json.load(f)
        except Exception:
            ncu = {}

        def tensor_roofline(launches, kernel):
            ms = sum(p["ms"] for p in launches)
            fl = sum(p["flops"] for p in launches)
            ach = fl / (ms * 1e-3) / 1e12
            return {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sust"], "frac_of_burst": ach / pk["tf_burst"],
                    "peak_source": pk["src"] + ", sustained bf16", "launches": len(launches),
                    "flops_per_launch": fl / len(launches), "us_per_launch": ms * 1e3 / len(launches),
                    "share_of_unet_time": ms / tot_ms,
                    "how": "algorithmic FLOPs (2*M*N*K of the GEMM as executed) / CUDA-event duration per launch, "
                           "one instrumented forward (events on the launching stream)"}

        roofline = tensor_roofline(top, "conv_igemm_kernel<BN=128, pair, halo, 2 M tiles/CTA> (3x3 convolutions at 256x256)")
        roofline["traffic"] = ncu.get("conv", {}).get("dram_bytes_per_launch")
        roofline["traffic_source"] = ncu.get("conv", {}).get("source")
        roofline["tensor_pipe_active_pct_ncu"] = ncu.get("conv", {}).get("tensor_pipe_active_pct_time_weighted")
        roofline_all = tensor_roofline(conv, "conv_igemm_kernel, all variants (every conv / linear / attention GEMM of the UNet)")
        roofline_all.update({"flops_per_forward": sum(p["flops"] for p in conv), "unet_ms": tot_ms,
                             "by_kind_ms": {k: round(v["ms"], 4) for k, v in by_kind.items()}})
        if args.profile_out:
            with open(args.profile_out, "w") as fo:
                json.dump(prof, fo)
        # ---- HBM roofline of the fused guided-step kernel at a batch that exceeds L2
        Bs = args.step_kernel_batch
        xs = torch.randn(Bs, 3, S, S, device=dev)
        es = torch.randn(Bs, 3, S, S, device=dev)
        zsb = torch.randn(3, S, S, device=dev)
        c = sch.coeffs(int(sch.timesteps[-5]), ETA, "ddpm")
        kw = dict(noise=zsb, targets=[TARGET, None, None], loss_scale=LOSS_SCALE, n_mean=S * S)
        for _ in range(3):
            ops.guided_step(xs, es, c, **kw)
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ops.guided_step(xs, es, c, **kw)
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / reps
        step_bytes = 16.0 * xs.numel()     # read x_t, eps; write x_prev, x0 (z is (C,H,W), amortised)
        step_gbs = step_bytes / (step_ms * 1e-3) / 1e9
        roofline_step = {"bound": "hbm", "kernel": "guided_step_vec4 (fused x0 + DDPM step + sigma*z + colour guidance)",
                         "achieved": step_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": step_gbs / pk["hbm"],
                         "peak_source": pk["src"], "traffic": ncu.get("step", {}).get("dram_bytes_per_launch"),
                         "traffic_source": ncu.get("step", {}).get("source"), "batch": Bs, "bytes_per_launch": step_bytes,
                         "ms_per_launch": step_ms,
                         "how": "16 B/elem algorithmic bytes / CUDA-event time, working set 4x%.0f MB > L2" % (xs.numel() * 4 / 1e6)}
        del xs, es
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, n, per = cpu_guided_steps(3, 1, batch=1, budget_s=25.0)
            cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"oracle port (torch-CPU fp32 UNet restatement + reference step math), batch 1, {n} steps, {per:.2f} s/step"}
        h2d = (x_host.numel() + zs_host[:K].numel()) * 4 / K
        d2h = (out_host.numel() + x0_host.numel()) * 4 / K
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "fp32-accurate (split bf16 hi+lo operands, 3 tensor-core products per GEMM, fp32 accumulation)",
                "data": "synthetic", "config": workload_config(args, B),
                "clocks": clk.summary(),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / K,
                        "how": "SegDiffEditPipeline.edit_image from pinned host x_T / z maps; the x0-prediction history "
                               "(copied step by step on a side stream, overlapped with the next UNet forward, joined before "
                               "the end event) and the final images land in pinned host memory inside the timed region"},
                "gpu_launches": launches, "roofline": roofline, "roofline_unet_convs": roofline_all,
                "roofline_step_kernel": roofline_step,
                "cpu_baseline": cpu,
                "unet_tflops_per_img": wrapper.unet.flops_per_sample / 1e12,   # as executed (fp32-accurate mode: 3x the algorithmic count)
                "unet_achieved_tflops": wrapper.unet.flops_per_sample * B / (tot_ms * 1e-3) / 1e12}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
