#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--batch B]

Workload ``pwcnet_train``: one self-supervised PWC-Net training step per "step" on synthetic brain-slice
pairs -- host model (cuDNN convolutions) around the ofr_b200 hot path: 5x Correlation(4,1,4,1,1) fwd+bwd,
4x pwc_warp fwd+bwd, the 7 stn warps (one batched launch each way), the fused 7-scale OFEloss fwd+bwd -- then Adam.  This is
BASELINE.json configs[2] (PWC-Net, batch 24 per GPU, 512x512) and, over N GPUs, configs[4] (DDP, one
NCCL gradient all-reduce per step).  Rank 0 prints ONE JSON line.

``--impl reference`` times the reference's CPU path instead: the same step with every hot-path layer
replaced by the CPU oracle (PyTorch CPU: shift-and-multiply correlation, F.grid_sample warps, loss.py
arithmetic) on the host cores, on a bounded sample (small batch) of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "PWC-Net train slice-pairs/s @1/2/4/8 B200; corr+warp kernel % of roofline"
UNIT = "slice-pairs/s"
FMA_PEAK_TFLOPS = 72.3  # measured: profiles/r01_fma_peak.txt
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of tools/opbench.py at the bench's
# shapes (B = 24, S = 512); keyed by the label of the C-ABI call
NCU_TRAFFIC_BYTES = {
    "corr_bwd C32 128x128 D2=81 g2": (228.1e6 + 77.4e6, "profiles/r01_u_ncu_corr_bwd.txt"),
    "corr_fwd C32 128x128 D9": (100.8e6 + 73.7e6, "profiles/r01_o_ncu_corr_fwd.txt"),
    "warp_bwd m2 C32 128x128 gI1=1 gI2=1": (107.0e6 + 30.9e6 + 3.2e6, "profiles/r01_t_ncu_corr_bwd_warp_bwd.txt"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ model builders
def build_gpu_step(device, channels_last=False, fused_loss=True):
    import ofr_b200 as ofr
    from ofr_b200.harness import PWCNetHost, RegistrationStep
    torch.manual_seed(6)
    host = PWCNetHost(md=4).to(device)
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
    for _ in range(warmup):
        train_step(step, opt, pair)
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        train_step(step, opt, pair)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    sec = sum(times) / len(times)
    return batch / sec, sec, len(times), torch.get_num_threads()


# ------------------------------------------------------------------------------------------------ main arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    batch = args.cpu_batch
    value, sec, done, threads = time_cpu(batch, args.size, args.steps, args.warmup)
    sample = (f"{done} steps of batch {batch} (of the GPU arm's {args.batch}) at {args.size}x{args.size}, "
              f"full training step on the CPU oracle layers")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"pwcnet_train_b{args.batch}_{args.size}x{args.size}", "cpu_sample_batch": batch,
                   "host_model": "PWCNetHost (PWC-DC-Net 1ch, 9374184 params) + oracle layers",
                   "optimizer": "Adam(lr=1e-4, eps=1e-4)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from ofr_b200.harness import synthetic_pairs, make_optimizer, train_step
    from ofr_b200.profiling import KernelTimer

    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    step, ofr = build_gpu_step(device, channels_last=args.channels_last, fused_loss=bool(args.fused_loss))
    if world > 1:
        step.predictor = torch.nn.parallel.DistributedDataParallel(step.predictor, device_ids=[local],
                                                                   gradient_as_bucket_view=True)
    opt = make_optimizer(step.predictor)
    B, S = args.batch, args.size
    pair_host = synthetic_pairs(B, S, seed=6 + rank).pin_memory()
    pair_dev = pair_host.to(device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        train_step(step, opt, pair_dev)

    # ---- timed region 1: inputs resident in HBM
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = ofr.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with KernelTimer() as kt:
        e0.record()
        for _ in range(args.steps):
            train_step(step, opt, pair_dev)
        e1.record()
    barrier()
    launches = ofr.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    ln = torch.tensor([float(launches)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    ms_per_step = ms.item() / args.steps
    value = world * B / (ms_per_step * 1e-3)
    ksum = kt.summary()

    # ---- timed region 2: end to end (pinned host batch -> device each step, loss read back each step)
    barrier()
    e0.record()
    last = None
    for _ in range(args.steps):
        pair = pair_host.to(device, non_blocking=True)
        last = train_step(step, opt, pair).item()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (ms2.item() / args.steps * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        rows = []
        for label, r in ksum.items():
            rows.append({"kernel": label, "calls_per_step": r["calls"] / args.steps, "avg_us": round(r["avg_us"], 2),
                         "algorithmic_bytes": r["bytes"], "gbs": round(r["gbs"], 1), "hbm_frac": round(r["gbs"] / peak, 4),
                         "tflops": round(r["tflops"], 3)})
        rows.sort(key=lambda r: -r["avg_us"] * r["calls_per_step"])
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
                    "hot_path_ms_per_step": round(ours_ms, 3), "hot_path_share_of_step": round(ours_ms / ms_per_step, 4)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, sec, done, threads = time_cpu(args.cpu_batch, S, steps=3, warmup=1, budget_s=25.0)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{done} steps of batch {args.cpu_batch} (of {B}) at {S}x{S}, same training step on the "
                             f"CPU oracle layers, {sec:.2f} s/step"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"pwcnet_train_b{B}_{S}x{S}", "per_gpu_batch": B, "global_batch": world * B,
                       "host_model": "PWCNetHost (PWC-DC-Net 1ch, 9374184 params; cuDNN convs, TF32 conv default)",
                       "optimizer": "Adam(lr=1e-4, eps=1e-4)", "parallelism": f"dp{world}",
                       "host_memory_format": "channels_last" if args.channels_last else "contiguous (NCHW)",
                       "cudnn_benchmark": bool(args.cudnn_benchmark),
                       "loss": "fused stn + OFEloss (ofr_regloss_fwd/bwd)" if args.fused_loss else "stn_multi then OFEloss",
                       "l2": "working set (activations > 1 GB per step) exceeds the 126 MB L2; no explicit flush"},
            "roofline": roofline, "kernels": rows[:24], "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pair_host.numel() * 4 * world,
                    "d2h_bytes_per_step": 8 * world, "last_loss": last},
            "gpu_launches": int(ln.item()), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512, help="slice side (BASELINE configs[2]: 512)")
    ap.add_argument("--batch", type=int, default=24, help="slice pairs per GPU (train.py default batch_size 24)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="bounded CPU sample: pairs per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--channels-last", type=int, default=0, help="host model (cuDNN convolutions) in channels_last")
    ap.add_argument("--fused-loss", type=int, default=1,
                    help="1: stn warps + OFEloss as the fused registration_loss; 0: stn_warp_multi then OFEloss")
    ap.add_argument("--cudnn-benchmark", type=int, default=1, help="torch.backends.cudnn.benchmark for the host model")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
