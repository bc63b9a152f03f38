#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see the task statement / DESIGN.md section 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--batch B]
                    [--scaling weak|strong --global-batch G] [--cuda-graph 0|1] [--secondary 0|1]

Headline workload ``pwcnet_train``: one self-supervised PWC-Net training step per "step" on synthetic brain-slice
pairs -- host model (cuDNN convolutions) around the ofr_b200 hot path: 5x Correlation(4,1,4,1,1) fwd+bwd, 4x pwc_warp
fwd+bwd, the fused registration loss (7 stn warps + 7-scale OFEloss, one launch each way) -- then Adam.  This is
BASELINE.json configs[2] (PWC-Net, batch 24 per GPU, 512x512) and, over N GPUs, configs[4] (DDP, one NCCL gradient
all-reduce per step; ``--scaling strong --global-batch 192`` splits a fixed global batch instead).  The step is timed
eagerly and, with ``--cuda-graph 1`` (default), as ONE captured CUDA graph (forward, loss, backward, Adam); the
faster of the two is ``value``.  Per-kernel rows come from a separate instrumented pass, never from the timed loops.

Secondary workload (N = 1 only, after the timed regions): a FlowNet2 training step (BASELINE configs[1],[3]) through
the REFERENCE'S OWN ``flownet2/models.py`` class, imported unmodified from ``baseline/_ref/`` with the ofr_b200 layers
at its import paths: Correlation(20,1,20,1,2) at C=256 (tcgen05), 4x Resample2d, 6x ChannelNorm, fwd + bwd.

``--impl reference`` times the reference's CPU path instead: the same PWC-Net step with every hot-path layer served by
the CPU oracle (PyTorch CPU: shift-and-multiply correlation, F.grid_sample warps, loss.py arithmetic) on the host
cores at the SAME batch and size, bounded in steps.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "PWC-Net train slice-pairs/s @1/2/4/8 B200; corr+warp kernel % of roofline"
UNIT = "slice-pairs/s"
FMA_PEAK_TFLOPS = 72.3    # measured: profiles/r01_fma_peak.txt (tools/micro/fma_peak.cu)
TF32_MMA_PEAK_TFLOPS = 1096.0  # measured: profiles/r01_tf32_umma_micro.txt (tcgen05.mma kind::tf32, M=128, N>=128)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of tools/opbench.py at the bench's
# shapes (B = 24, S = 512); keyed by the label of the C-ABI call
NCU_TRAFFIC_BYTES = {
    "corr_bwd C32 128x128 D2=81 g2": (228.1e6 + 75.1e6, "profiles/r02_ncu_corr_bwd_box.txt"),
    # ofr_flowconv3_bwd = conv3_bwd_x_kernel + conv3_bwd_w_kernel + wgrad_sum_kernel (reads + writes of the three)
    "flowconv3_bwd C565 128x128": ((3.2 + 829.0 + 899.4 + 5.2 + 0.7) * 1e6, "profiles/r02_ncu_flow_heads.txt"),
    "flowconv3_fwd C565 128x128": ((889.0 + 13.5 + 12.6) * 1e6, "profiles/r02_ncu_flow_heads.txt"),
    "corr_fwd C32 128x128 D9": (100.8e6 + 73.7e6, "profiles/r01_o_ncu_corr_fwd.txt"),
    "warp_bwd m2 C32 128x128 gI1=1 gI2=1": (107.0e6 + 30.9e6 + 3.2e6, "profiles/r01_t_ncu_corr_bwd_warp_bwd.txt"),
}
try:  # newer captures, written by tools/ncu_traffic.py from this round's .ncu-rep files
    NCU_TRAFFIC_BYTES.update({k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).items()})
except (OSError, ValueError):
    pass


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ model builders
def build_gpu_step(device, channels_last=False, fused_loss=True, flow_heads=True):
    import ofr_b200 as ofr
    from ofr_b200.harness import PWCNetHost, RegistrationStep
    torch.manual_seed(6)
    # flow_heads: predict_flow* / dc_conv7 / deconv* / upfeat* (two output channels) on ofr_flowconv3 / ofr_flowdeconv4
    # instead of cuDNN; same parameters
    host = PWCNetHost(md=4, fast_heads=flow_heads).to(device)
    if channels_last:  # host-model memory format only; the ofr_b200 layers take / return NCHW (they call .contiguous())
        host = host.to(memory_format=torch.channels_last)
    host.deconv0.requires_grad_(False)  # unused in the forward (PWCNet.py:126,274): keep it out of DDP's reducer
    # fused_loss: the stn warps of all scales + OFEloss as one launch each way (ofr_regloss_fwd/bwd, bit-identical to
    # stn_warp_multi + OFEloss)
    return RegistrationStep(host, ofr.stn_warp, ofr.OFEloss, stn_multi=ofr.stn_warp_multi,
                            reg_loss=ofr.registration_loss if fused_loss else None), ofr


def build_cpu_step():
    """The reference's CPU path: same host model, every hot-path layer served by the oracle."""
    import oracle
    from ofr_b200.harness import PWCNetHost, RegistrationStep

    class OracleCorrelation(torch.nn.Module):
        def forward(self, a, b):
            return oracle.correlation(a, b, 4, 1, 4, 1, 1)

    torch.manual_seed(6)
    host = PWCNetHost(md=4, corr=OracleCorrelation(), warp=oracle.pwc_warp)
    host.deconv0.requires_grad_(False)
    return RegistrationStep(host, oracle.stn_warp, oracle.ofe_loss)


def time_cpu(batch, size, steps, warmup, budget_s=None):
    from ofr_b200.harness import synthetic_pairs, make_optimizer, train_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = build_cpu_step()
    opt = make_optimizer(step.predictor)
    pair = synthetic_pairs(batch, size, seed=6)
    t_begin = time.perf_counter()
    for _ in range(warmup):
        train_step(step, opt, pair)
        if budget_s is not None and time.perf_counter() - t_begin > 0.4 * budget_s:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        train_step(step, opt, pair)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    sec = sum(times) / len(times)
    return batch / sec, sec, len(times), torch.get_num_threads()


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    batch = args.cpu_batch if args.cpu_batch > 0 else args.batch
    # bounded: the whole run ends within a few minutes whatever --steps asks for (a batch-24 512x512 step is ~7 s)
    value, sec, done, threads = time_cpu(batch, args.size, args.steps, min(args.warmup, 1), budget_s=150.0)
    sample = (f"{done} steps of batch {batch} (GPU arm: {args.batch}) at {args.size}x{args.size}, full training step "
              f"on the CPU oracle layers, {sec:.2f} s/step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"pwcnet_train_b{args.batch}_{args.size}x{args.size}", "cpu_sample_batch": batch,
                   "same_config": batch == args.batch,
                   "host_model": "PWCNetHost (PWC-DC-Net 1ch, 9374184 params) + oracle layers",
                   "optimizer": "Adam(lr=1e-4, eps=1e-4)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def kernel_rows(ksum, steps, peak):
    rows = []
    for label, r in ksum.items():
        row = {"kernel": label, "calls_per_step": r["calls"] / steps, "avg_us": round(r["avg_us"], 2),
               "algorithmic_bytes": r["bytes"], "gbs": round(r["gbs"], 1), "hbm_frac": round(r["gbs"] / peak, 4),
               "tflops": round(r["tflops"], 3)}
        if " D21" in label or "D2=441" in label:
            # tcgen05 3xTF32 banded-Gram kernels, reported against the measured tf32 MMA peak.  Executed MMA work per
            # useful MAC: forward 3 products x 4 chunks x N=256 neighbourhood columns per 441 displacements; backward
            # 3 products x (neighbourhood rows inside the image, <= 28) x K=40 columns per 441 displacements.
            hc = int(label.split()[2].split("x")[0]) // 2  # class-image height
            rows_walked = sum(min(28, hc + 10 - y0) - max(0, 10 - y0) for y0 in range(0, hc, 8)) / max(hc // 8, 1)
            ex = 3.0 * (4 * 256) / 441.0 if label.startswith("corr_fwd") else 3.0 * (rows_walked * 40) / 441.0
            row["bound"] = "tensor"
            row["mma_tflops_executed"] = round(r["tflops"] * ex, 1)
            row["tf32_mma_frac"] = round(r["tflops"] * ex / TF32_MMA_PEAK_TFLOPS, 4)
            row["fp32_equiv_frac_of_3xtf32_band_peak"] = round(r["tflops"] / (TF32_MMA_PEAK_TFLOPS / ex), 4)
        rows.append(row)
    rows.sort(key=lambda r: -r["avg_us"] * r["calls_per_step"])
    return rows


def run_secondary_flownet2(device, args, peak):
    """FlowNet2 training step through the reference's own class (baseline/_ref) on the ofr_b200 layers."""
    import ofr_b200 as ofr
    from ofr_b200.harness import reference_env, RegistrationStep, synthetic_pairs, make_optimizer, train_step
    from ofr_b200.profiling import KernelTimer
    if reference_env.reference_dir() is None:
        return {"unavailable": "baseline/_ref not staged (python tools/stage_reference.py in the dev container)"}
    B, S = args.fn2_batch, args.fn2_size
    with reference_env.ReferenceEnv("ofr") as e:
        torch.manual_seed(6)
        net = e.flownet2().to(device)
    heads = ofr.accelerate_flow_heads(net) if args.flow_heads else []   # predict_flow* / upsampled_flow* of the 5 sub-networks
    step = RegistrationStep(net, ofr.stn_warp, ofr.OFEloss, reg_loss=ofr.registration_loss)
    opt = make_optimizer(net)
    pair = synthetic_pairs(B, S, seed=7).to(device)
    for _ in range(3):
        train_step(step, opt, pair)
    torch.cuda.synchronize()
    n0 = ofr.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.fn2_steps):
        train_step(step, opt, pair)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.fn2_steps
    launches = (ofr.launch_count() - n0) / args.fn2_steps
    with KernelTimer() as kt:
        for _ in range(2):
            train_step(step, opt, pair)
    rows = kernel_rows(kt.summary(), 2, peak)
    ours_ms = sum(r["avg_us"] * r["calls_per_step"] for r in rows) * 1e-3
    dom = rows[0]
    roof = {"kernel": dom["kernel"], "avg_launch_us": dom["avg_us"], "algorithmic_bytes": dom["algorithmic_bytes"]}
    if dom.get("bound") == "tensor":
        roof.update(bound="tensor", achieved=dom["mma_tflops_executed"], peak=TF32_MMA_PEAK_TFLOPS, unit="TFLOP/s",
                    frac=dom["tf32_mma_frac"], useful_tflops_fp32_equiv=dom["tflops"],
                    peak_source="measured tcgen05 kind::tf32 MMA rate (tools/micro/tf32_umma.cu); `achieved` counts "
                                "the executed 3xTF32 banded-Gram MMAs, `useful` the operator's own FLOPs")
    else:
        roof.update(bound="hbm", achieved=dom["gbs"], peak=peak, unit="GB/s", frac=dom["hbm_frac"])
    del net, step, opt
    torch.cuda.empty_cache()
    return {"workload": f"flownet2_train_b{B}_{S}x{S}", "predictor": "reference flownet2/models.py FlowNet2 (unmodified, "
            "baseline/_ref) on ofr_b200 Correlation/Resample2d/ChannelNorm", "flow_heads_reclassed": len(heads),
            "value": B / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms, "steps": args.fn2_steps, "gpu_launches_per_step": launches,
            "hot_path_ms_per_step": round(ours_ms, 3), "roofline": roof, "kernels": rows}


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    use_graph = bool(args.cuda_graph)
    if world > 1:
        if use_graph:  # whole-step capture with DDP (PyTorch CUDA-graphs note): no async NCCL error handling
            os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        # NCCL prints its version banner (and, with NCCL_DEBUG set, its log) on stdout while the communicator is built:
        # point fd 1 at stderr for that moment, so that stdout carries the ONE JSON line and nothing else
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()  # first collective: the communicator exists after it
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from ofr_b200.harness import synthetic_pairs, make_optimizer, train_step, GraphedStep
    from ofr_b200.profiling import KernelTimer, ClockSampler

    if args.scaling == "strong":
        assert args.global_batch % world == 0, "--global-batch must divide by the number of GPUs"
        B = args.global_batch // world
    else:
        B = args.batch
    S = args.size
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    step, ofr = build_gpu_step(device, channels_last=args.channels_last, fused_loss=bool(args.fused_loss),
                               flow_heads=bool(args.flow_heads))
    if world > 1:
        ddp_kw = dict(device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=args.bucket_cap_mb,
                      static_graph=True)
        if use_graph:  # DDP must be constructed in a side-stream context for full-backward capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step.predictor = torch.nn.parallel.DistributedDataParallel(step.predictor, **ddp_kw)
            torch.cuda.current_stream().wait_stream(side)
        else:
            step.predictor = torch.nn.parallel.DistributedDataParallel(step.predictor, **ddp_kw)
    opt = make_optimizer(step.predictor, capturable=use_graph)
    pair_host = synthetic_pairs(B, S, seed=6 + rank).pin_memory()
    pair_dev = pair_host.to(device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_rank = {}

    def max_over_ranks(ms, tag=None):
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            if tag is not None:  # every rank's own device time for the region (the reported time is their maximum)
                every = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(every, t)
                per_rank[tag] = [round(float(e.item()), 3) for e in every]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(max(args.warmup, 11 if (use_graph and world > 1) else 0)):
        train_step(step, opt, pair_dev)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- timed region 1a: eager step, inputs resident in HBM (nothing instrumented inside the loop)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = ofr.launch_count()
    e0.record()
    for _ in range(args.steps):
        train_step(step, opt, pair_dev)
    e1.record()
    barrier()
    launches = ofr.launch_count() - n0
    eager_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, "eager_ms_per_step")

    # ---- timed region 1b: the same step as one CUDA graph
    graphed, graph_ms, graph_err = None, None, None
    if use_graph:
        try:
            graphed = GraphedStep(step, opt, pair_dev, warm_iters=11 if world > 1 else 3)
            for _ in range(args.warmup):
                graphed.replay()
            barrier()
            e0.record()
            for _ in range(args.steps):
                graphed.replay()
            e1.record()
            barrier()
            graph_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, "cuda_graph_ms_per_step")
        except Exception as ex:  # capture refused (it never silently changes the workload): report and keep eager
            graphed, graph_err = None, f"{type(ex).__name__}: {str(ex)[:200]}"
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    headline_graph = graph_ms is not None and graph_ms <= eager_ms
    ms_per_step = graph_ms if headline_graph else eager_ms
    value = world * B / (ms_per_step * 1e-3)
    ln = torch.tensor([float(launches)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)

    # ---- timed region 2: end to end (pinned host batch -> device each step, loss read back each step).  The copy of
    # step i+1 runs on a copy stream while step i computes (double-buffered staging); every step still moves its own
    # batch host -> device inside the timed region and reads its loss back.
    copy_stream = torch.cuda.Stream()
    stage = [torch.empty_like(pair_dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def start_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])   # the step that read this staging buffer has taken its data
            stage[i % 2].copy_(pair_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    for ev in consumed:
        ev.record()
    barrier()
    e0.record()
    last = None
    start_copy(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            start_copy(i + 1)
        torch.cuda.current_stream().wait_event(ready[i % 2])
        if headline_graph:
            graphed.static_in.copy_(stage[i % 2], non_blocking=True)
            consumed[i % 2].record()
            last = graphed.replay().item()
        else:
            pair = stage[i % 2]
            loss_t = train_step(step, opt, pair)
            consumed[i % 2].record()
            last = loss_t.item()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = world * B / (e2e_ms * 1e-3)

    # ---- instrumented pass (outside every timed region): CUDA events around each C-ABI call
    ksteps = 3
    with KernelTimer() as kt:
        for _ in range(ksteps):
            train_step(step, opt, pair_dev)
    ksum = kt.summary()
    barrier()

    if rank == 0:
        peak, peak_src = peaks()
        rows = kernel_rows(ksum, ksteps, peak)
        ours_ms = sum(r["avg_us"] * r["calls_per_step"] for r in rows) * 1e-3
        dom = rows[0]
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(dom["gbs"] / peak, 4),
                    "traffic": NCU_TRAFFIC_BYTES.get(dom["kernel"], (None, None))[0],
                    "traffic_source": NCU_TRAFFIC_BYTES.get(dom["kernel"], (None, None))[1],
                    "algorithmic_bytes": dom["algorithmic_bytes"], "peak_source": peak_src,
                    "avg_launch_us": dom["avg_us"], "tflops_fp32": dom["tflops"],
                    "fp32_fma_peak_tflops": FMA_PEAK_TFLOPS, "fp32_fma_frac": round(dom["tflops"] / FMA_PEAK_TFLOPS, 4),
                    "fp32_fma_peak_source": "measured on this pool's B200 with tools/micro/fma_peak.cu (d = a*b + d, "
                                            "register operands; nominal 148 SM x 128 x 2 x 1.965 GHz = 74.5)",
                    "hot_path_ms_per_step": round(ours_ms, 3), "hot_path_share_of_step": round(ours_ms / eager_ms, 4),
                    "hot_path_algorithmic_bytes": sum(r["algorithmic_bytes"] * r["calls_per_step"] for r in rows),
                    "hot_path_hbm_frac": round(sum(r["algorithmic_bytes"] * r["calls_per_step"] for r in rows)
                                               / (ours_ms * 1e-3) / 1e9 / peak, 4) if ours_ms > 0 else None}
        if dom["kernel"].startswith("flow"):
            # a flow-head row is one C-ABI call, i.e. up to three launches; the per-kernel ncu figures are in the profile
            roofline["note"] = ("row = one ofr_flow*_{fwd,bwd} call (forward: channel-split kernel + split sum; backward: "
                                "input-gradient kernel + weight-gradient kernel + split sum); per-kernel durations and "
                                "DRAM bytes: profiles/r02_ncu_flow_heads.txt")
        secondary = None
        if world == 1 and args.secondary:
            try:
                secondary = run_secondary_flownet2(device, args, peak)
            except Exception as ex:
                secondary = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb = args.cpu_batch if args.cpu_batch > 0 else B
            v, sec, done, threads = time_cpu(cb, S, steps=3, warmup=1, budget_s=40.0)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{done} steps of batch {cb} (GPU arm: {B}) at {S}x{S}, same training step on the "
                             f"CPU oracle layers, {sec:.2f} s/step"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"pwcnet_train_b{B}_{S}x{S}", "per_gpu_batch": B, "global_batch": world * B,
                       "host_model": "PWCNetHost (PWC-DC-Net 1ch, 9374184 params; cuDNN convs, TF32 conv default)",
                       "optimizer": "Adam(lr=1e-4, eps=1e-4" + (", capturable)" if use_graph else ")"),
                       "parallelism": f"dp{world}", "cuda_graph": bool(headline_graph),
                       "host_memory_format": "channels_last" if args.channels_last else "contiguous (NCHW)",
                       "cudnn_benchmark": bool(args.cudnn_benchmark), "ddp_bucket_cap_mb": args.bucket_cap_mb if world > 1 else None,
                       "loss": "fused stn + OFEloss (ofr_regloss_fwd/bwd)" if args.fused_loss else "stn_multi then OFEloss",
                       "flow_heads": "ofr_flowconv3 / ofr_flowdeconv4 (fp32)" if args.flow_heads else "cuDNN",
                       "l2": "working set (activations > 1 GB per step) exceeds the 126 MB L2; no explicit flush"},
            "step_modes": {"eager_ms_per_step": eager_ms, "cuda_graph_ms_per_step": graph_ms, "cuda_graph_error": graph_err,
                           "headline": "cuda_graph" if headline_graph else "eager",
                           "per_rank_ms_per_step": per_rank or None},
            "roofline": roofline, "kernels": rows[:24], "secondary": secondary, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pair_host.numel() * 4 * world,
                    "d2h_bytes_per_step": 8 * world, "last_loss": last, "ms_per_step": e2e_ms,
                    "h2d": "pinned -> device on a copy stream, double-buffered: the copy of step i+1 overlaps step i"},
            "gpu_launches": int(ln.item()), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing the communicator down: destroying it while a captured graph still references its
        # collectives hung the process at exit (measured at N = 2); the JSON line is out, every rank is past its last
        # collective
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        del graphed
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512, help="slice side (BASELINE configs[2]: 512)")
    ap.add_argument("--batch", type=int, default=24, help="slice pairs per GPU (train.py default batch_size 24)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch pairs per GPU; strong: --global-batch split over the GPUs (BASELINE configs[4])")
    ap.add_argument("--global-batch", type=int, default=192)
    ap.add_argument("--cpu-batch", type=int, default=0, help="pairs per CPU step of the reference arm (0 = same as the GPU arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--channels-last", type=int, default=0, help="host model (cuDNN convolutions) in channels_last")
    ap.add_argument("--fused-loss", type=int, default=1,
                    help="1: stn warps + OFEloss as the fused registration_loss; 0: stn_warp_multi then OFEloss")
    ap.add_argument("--flow-heads", type=int, default=1,
                    help="1: the decoder's two-output-channel layers on the flow-head kernels (SURVEY 8f-4); 0: cuDNN")
    ap.add_argument("--cudnn-benchmark", type=int, default=1, help="torch.backends.cudnn.benchmark for the host model")
    ap.add_argument("--cuda-graph", type=int, default=1, help="also time the step as one captured CUDA graph")
    ap.add_argument("--bucket-cap-mb", type=float, default=8.0,
                    help="DDP gradient bucket size (measured at N = 8: 8 MB 57.7 ms, 25 MB 58.1, one 100 MB bucket 58.2)")
    ap.add_argument("--secondary", type=int, default=1, help="run the FlowNet2 secondary workload (N = 1)")
    ap.add_argument("--fn2-size", type=int, default=512)
    ap.add_argument("--fn2-batch", type=int, default=24)
    ap.add_argument("--fn2-steps", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
