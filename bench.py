#!/usr/bin/env python
"""Benchmark of the MIRAGE MultiViT hot path on B200 (the driver's contract).

  python bench.py --gpus N --steps K --warmup W [--workload encoder_large|pretrain_large|...]
  python bench.py --impl reference ...     # the reference's CPU arithmetic (oracle port) on host cores

Default workload (N=1) = BASELINE.json configs[1]: MIRAGE-Large encoder inference, bscan+slo 512x512,
256 images per GPU per step, bf16 tensor-core math.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# algorithmic FLOPs per sample (forward), SURVEY.md section 8(d) / BASELINE.md section 4
GFLOP_FWD = {"encoder_base": 97.65, "encoder_large": 336.79, "pretrain_base": 24.07, "pretrain_large": 68.49,
             "cls_large": 162.25}

# of which: input-adapter patch projections of ALL tokens (column "patch-embed")
EMBED_GFLOP_FWD = {"pretrain_base": 2.42, "pretrain_large": 3.22}

WORKLOADS = {
    # name: (size, modalities, per-GPU batch, kind)
    "encoder_large": ("large", ["bscan", "slo"], 256, "encoder"),
    "encoder_base": ("base", ["bscan", "slo"], 256, "encoder"),
    "pretrain_large": ("large", ["bscan", "slo", "bscanlayermap"], 256, "pretrain"),
    "pretrain_base": ("base", ["bscan", "slo", "bscanlayermap"], 256, "pretrain"),
    "cls_large": ("large", ["bscan"], 64, "cls"),   # BASELINE configs[4]: fine-tune fwd+bwd, 64 per GPU
}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# oracle (CPU) legs
# ---------------------------------------------------------------------------------------------
def cpu_oracle_throughput(workload: str, budget_s: float = 20.0, batch: int = 2, seed: int = 0):
    """Times the reference's arithmetic (oracle port, fp32, all host threads) on a bounded sample of the
    same workload.  Returns (samples/s, cores, description)."""
    from oracle import mirage_oracle as O
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import synth_images, synth_state_dict
    size, mods, _, kind = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    depth, heads = (12, 12) if size == "base" else (24, 16)
    if kind == "encoder":
        # shapes from the oracle itself: the CPU arm never imports the product package
        sd = O.light_state_dict_shapes(mods, size)
        sd.update(synth_state_dict({k: v.shape for k, v in sd.items()}, seed))
        x = synth_images(batch, mods, seed=1234)

        def run():
            with torch.no_grad():
                return O.light_forward(x, sd, depth, heads)
    elif kind == "cls":
        from bench_support import build_cls_oracle
        run = build_cls_oracle(size, batch, seed)
    else:
        from bench_support import build_pretrain_oracle
        run = build_pretrain_oracle(size, mods, batch, seed)
    run()  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        run()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 8:
            break
    return batch * n / el, cores, f"{n} x batch {batch} of {workload}, fp32 oracle, {cores} threads"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, args.steps), args.warmup
    budget = min(25.0, 120.0 / max(1, steps))
    vals = []
    desc = ""
    cores = os.cpu_count() or 1
    for i in range(min(warm, 1) + steps):
        v, cores, desc = cpu_oracle_throughput(args.workload, budget_s=budget / 2, batch=2)
        if i >= min(warm, 1):
            vals.append(v)
    value = sum(vals) / len(vals)
    size, mods, per_gpu, kind = WORKLOADS[args.workload]
    unit = "images/s" if kind == "encoder" else "samples/s"
    line = {
        "impl": "reference", "metric": f"{args.workload} {unit}", "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * 2 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "modalities": mods, "per_step_sample": desc},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class KernelTimer:
    """Per-launch CUDA-event timing on the launching stream, grouped by kernel name."""

    def __init__(self):
        self.records = []

    def __call__(self, name, work, unit):
        timer = self

        class _Ctx:
            def __enter__(self_):
                self_.e0 = torch.cuda.Event(enable_timing=True)
                self_.e1 = torch.cuda.Event(enable_timing=True)
                self_.e0.record()

            def __exit__(self_, *a):
                self_.e1.record()
                timer.records.append((name, work, unit, self_.e0, self_.e1))
                return False
        return _Ctx()

    def summary(self):
        agg = {}
        for name, work, unit, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            a = agg.setdefault(name, {"ms": 0.0, "work": 0.0, "unit": unit, "launches": 0})
            a["ms"] += ms
            a["work"] += work
            a["launches"] += 1
        return agg


def _sync_all(world):
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def _timed(step, K, world, dev):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events, max over ranks (ms total)."""
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _sync_all(world)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    _sync_all(world)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _traffic_entry(workload, kernel):
    """DRAM bytes per launch of `kernel` in `workload` from the committed ncu --set full captures."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tr = json.loads((ROOT / "profiles" / name).read_text())
            v = tr.get(workload, {}).get(kernel, {}).get("bytes_per_launch")
            if v is not None:
                return v
        except Exception:
            continue
    return None


def measure_workload(args, workload, per_gpu, dev, rank, world, local, sample_clocks=True):
    """One workload, whole protocol: build, (graph) capture, W warm-up steps, K timed steps, K end-to-end
    steps, one instrumented step for the per-kernel roofline.  Returns the record (rank 0) or None."""
    from mirage_b200 import ops
    size, mods, _, kind = WORKLOADS[workload]
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import load_synth, synth_images
    extra = {}
    graph_hooks = None
    ddp = None
    if kind == "encoder":
        from mirage_b200.mirage_hf import MIRAGEWrapper
        model = MIRAGEWrapper(size=size, modalities="-".join(mods))
        load_synth(model.model, seed=0)
        model = model.to(dev).eval()
        base = synth_images(8, mods, seed=1234 + rank)
        host_in = {k: v.repeat(per_gpu // 8 + 1, 1, 1, 1)[:per_gpu].contiguous().pin_memory()
                   for k, v in base.items()}
        dev_in = {k: v.to(dev) for k, v in host_in.items()}
        n_tok = sum((v.shape[-1] // 32) * (v.shape[-2] // 32) for v in dev_in.values()) + 1
        D = model.model.dim_tokens
        host_out = torch.empty((per_gpu, n_tok, D), dtype=torch.float32).pin_memory()

        def step_eager():
            with torch.no_grad():
                return model(dev_in)

        def step_e2e():
            # public host-to-host call: pinned inputs -> encoder -> pinned output, copies overlapped
            # with compute chunk by chunk (mirage_hf.MIRAGEWrapper.encode_host)
            return model.encode_host(host_in, out=host_out, chunk=args.e2e_chunk,
                                     ramp=min(args.e2e_ramp, max(1, per_gpu // 8)))
        h2d = sum(v.numel() * v.element_size() for v in host_in.values())
        d2h = host_out.numel() * host_out.element_size()
        flop_per_sample = GFLOP_FWD[workload] * 1e9
        unit = "images/s"
        # the same call fed with the RAW uint8 images (the on-disk format, scaled to [0, 1] on the device)
        host_u8 = {k: (v * 255.0).round().to(torch.uint8).pin_memory() for k, v in host_in.items()}

        def step_e2e_u8():
            return model.encode_host(host_u8, out=host_out, chunk=args.e2e_chunk,
                                     ramp=min(args.e2e_ramp, max(1, per_gpu // 8)))
        extra["_e2e_u8"] = (step_e2e_u8, sum(v.numel() for v in host_u8.values()))
    elif kind == "cls":
        from bench_support import build_cls_step
        step_eager, step_e2e, h2d, d2h, graph_hooks, ddp = build_cls_step(size, per_gpu, dev, rank, world, args)
        flop_per_sample = 3.0 * GFLOP_FWD[workload] * 1e9
        unit = "samples/s"
    else:
        from bench_support import build_pretrain_step
        step_eager, step_e2e, h2d, d2h, graph_hooks, ddp = build_pretrain_step(size, mods, per_gpu, dev, rank,
                                                                               world, args)
        flop_per_sample = 3.0 * GFLOP_FWD[workload] * 1e9
        # The masked forward embeds only the kept patches (mirage_b200/csrc/visible.cu) where the reference embeds
        # all 768 and gathers 98: that work is NOT done here, so it is not counted either -- model_tflops uses
        # the executed count, and the reference's full count is reported next to it.
        if os.environ.get("MB_EMBED_VISIBLE", "1") != "0":
            extra["model_gflop_per_sample_reference_count"] = round(flop_per_sample / 1e9, 2)
            flop_per_sample -= 3.0 * EMBED_GFLOP_FWD[workload] * 1e9 * (1.0 - 98.0 / 768.0)
            extra["embedding"] = "kept tokens only (98 of 768 patches per sample embedded)"
        unit = "samples/s"
    step = step_eager

    graphed = False
    if args.graph:
        try:
            if kind == "encoder":
                from mirage_b200.graphs import GraphedCallable
                step = GraphedCallable(step_eager, refresh_weights=False).capture()
                graphed = True
            elif graph_hooks is not None:
                step, step_e2e = graph_hooks()
                graphed = True
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"bench.py: CUDA-graph capture of {workload} failed ({type(e).__name__}: {e}); "
                      "eager launches", file=sys.stderr, flush=True)
            torch.cuda.synchronize()
            step = step_eager

    W, K = max(3, args.warmup), max(1, args.steps)
    for _ in range(W):
        step()
    sampler = ClockSampler(local) if (rank == 0 and sample_clocks) else None
    if sampler:
        sampler.start()
    ms_total = _timed(step, K, world, dev)
    clocks = sampler.stop() if sampler else None

    for _ in range(2):
        step_e2e()
    ms_e2e = _timed(step_e2e, K, world, dev)
    e2e_u8 = None
    if "_e2e_u8" in extra:
        fn, nbytes = extra.pop("_e2e_u8")
        for _ in range(2):
            fn()
        e2e_u8 = (_timed(fn, K, world, dev), nbytes)

    # how much of the gradient exchange is NOT hidden behind backward: the same step with the all-reduce off
    if ddp is not None and world > 1:
        ddp.enabled = False
        try:
            quiet = step_eager
            if graphed and graph_hooks is not None and args.mask_sampler == "device":
                quiet, _ = graph_hooks()
            for _ in range(2):
                quiet()
            ms_quiet = _timed(quiet, K, world, dev)
            extra["allreduce_exposed_ms"] = round((ms_total - ms_quiet) / K, 3)
            extra["ms_per_step_without_allreduce"] = round(ms_quiet / K, 3)
        finally:
            ddp.enabled = True
        extra["allreduce_bytes_per_step"] = int(sum(b.flat.numel() * b.flat.element_size() for b in ddp.buckets))
        extra["reserve_sms"] = ddp.reserve_sms
    elif ddp is not None:
        extra["allreduce_exposed_ms"] = 0.0

    # per-kernel roofline pass: one more (eager) step with CUDA events around every launch; every rank runs it
    # (the training step contains the collective), rank 0 records
    kt = KernelTimer()
    if rank == 0:
        ops.set_recorder(kt)
    ops.reset_launch_count()
    step_eager()
    torch.cuda.synchronize()
    ops.set_recorder(None)
    launches = ops.launch_count() * K
    _sync_all(world)
    if rank != 0:
        return None

    peaks = measured_peaks()
    agg = kt.summary()
    tot = sum(a["ms"] for a in agg.values()) or 1.0
    kernels = {k: {"ms": round(a["ms"], 3), "share": round(a["ms"] / tot, 3), "launches": a["launches"],
                   ("tflops" if a["unit"] == "flop" else "gbs"):
                       round(a["work"] / a["ms"] / (1e9 if a["unit"] == "flop" else 1e6), 1)}
               for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    name, a = max(agg.items(), key=lambda kv: kv[1]["ms"])
    traffic = _traffic_entry(workload, name)
    if a["unit"] == "flop":
        ach = a["work"] / a["ms"] / 1e9
        peak = peaks["bf16_tflops_sustained"]
        roofline = {"kernel": name, "bound": "tensor", "achieved": round(ach, 1), "peak": peak,
                    "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic,
                    "peak_source": peaks["source"] + " (sustained bf16)",
                    "per_launch": {"flops": a["work"] / a["launches"], "ms": a["ms"] / a["launches"]}}
    else:
        ach = a["work"] / a["ms"] / 1e6
        peak = peaks["hbm_gbs"]
        roofline = {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peak,
                    "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic,
                    "peak_source": peaks["source"],
                    "per_launch": {"bytes": a["work"] / a["launches"], "ms": a["ms"] / a["launches"]}}
    total = per_gpu * world * K
    value = total / (ms_total / 1e3)
    rec = {
        "metric": f"{workload} {unit}", "value": round(value, 2), "unit": unit, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": round(ms_total / K, 3),
        "config": {"workload": workload, "size": size, "modalities": mods, "per_gpu_batch": per_gpu,
                   "global_batch": per_gpu * world,
                   "parallelism": f"dp{world} (batch-sharded, no collective)" if kind == "encoder"
                   else f"dp{world} (NCCL gradient all-reduce overlapped with backward)",
                   "l2_policy": "inputs_exceed_l2 (activations >> 126 MB per step)", "cuda_graph": graphed},
        "model_tflops": round(value * flop_per_sample / 1e12, 1),
        "model_frac_of_bf16_peak": round(value * flop_per_sample / 1e12 / world / peaks["bf16_tflops"], 4),
        "roofline": roofline, "kernels": kernels,
        "e2e": {"value": round(total / (ms_e2e / 1e3), 2), "unit": unit, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "pcie_gbs_per_rank": round((h2d + d2h) / (ms_e2e / K / 1e3) / 1e9, 2)},
        "gpu_launches": launches, "clocks": clocks,
    }
    if e2e_u8 is not None:
        ms_u8, nbytes = e2e_u8
        rec["e2e_uint8_inputs"] = {"value": round(total / (ms_u8 / 1e3), 2), "unit": unit,
                                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h,
                                   "pcie_gbs_per_rank": round((nbytes + d2h) / (ms_u8 / K / 1e3) / 1e9, 2),
                                   "note": "same public call, inputs as raw uint8 (scaled to [0,1] on the device)"}
    if kind != "encoder":
        rec["config"]["optimizer"] = args.optimizer
        rec["config"]["mask_sampler"] = args.mask_sampler if kind == "pretrain" else None
    rec.update(extra)
    return rec


def _bind_to_gpu_numa(local: int):
    """Pin this rank to the CPU cores NVML reports as local to its GPU BEFORE any pinned buffer is allocated
    (first touch then places the staging memory on that NUMA node): with every rank on node 0 the eight
    host<->device streams of an 8-GPU run share one memory controller (SCALE_r01: e2e efficiency 0.889 at N=8).
    Returns a short description for the JSON line; never fails the run."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [i for i in range(n_cpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"gpu": local, "cpus": f"{allowed[0]}-{allowed[-1]}", "n_cpus": len(allowed)}
    except Exception as e:  # noqa: BLE001
        return {"gpu": local, "cpus": None, "note": f"{type(e).__name__}"}
    return {"gpu": local, "cpus": None}


def _free_gpu():
    import gc
    from mirage_b200 import functional as Fn
    Fn.set_grad_sink(None)
    Fn.drop_cast_cache()
    gc.collect()
    torch.cuda.empty_cache()


def run_gpu(args):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the MIRAGE hot path has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_binding = _bind_to_gpu_numa(local)
    if world > 1:
        if args.nccl_max_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_max_ctas))
        # NCCL announces itself on stdout ("NCCL version ...") while the communicator is created; stdout carries
        # exactly ONE JSON line, so file descriptor 1 points at stderr for the duration of the eager initialisation
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    # default invocation = both halves of BASELINE.json's metric: the MIRAGE-L encoder (headline line) and
    # the MultiMAE pretraining step (sub-record "pretrain"), plus the strong-scaling encoder point of SURVEY 8(d)
    # (global batch 256 split over the ranks) when there is more than one rank
    default = args.workload == "default"
    head_name = "encoder_large" if default else args.workload
    per_gpu = args.batch or WORKLOADS[head_name][2]
    head = measure_workload(args, head_name, per_gpu, dev, rank, world, local)
    subs = {}
    if default:
        if world > 1 and not args.skip_strong:
            _free_gpu()
            subs["encoder_strong"] = measure_workload(args, "encoder_large", max(1, 256 // world), dev, rank, world,
                                                      local, sample_clocks=False)
        if not args.skip_pretrain:
            _free_gpu()
            subs["pretrain"] = measure_workload(args, "pretrain_large", args.batch or WORKLOADS["pretrain_large"][2],
                                                dev, rank, world, local)

    if rank == 0:
        cpu = None
        unit = head["unit"]
        if world == 1 and not args.no_cpu_baseline:
            v, cores, desc = cpu_oracle_throughput(head_name, budget_s=15.0, batch=2)
            cpu = {"value": round(v, 3), "unit": unit, "cores": cores, "kind": "port", "sample": desc}
        line = {
            "metric": head["metric"], "value": head["value"], "unit": unit, "n_gpus": world,
            "steps": head["steps"], "warmup": head["warmup"], "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": head["config"],
            "model_tflops": head["model_tflops"], "model_frac_of_bf16_peak": head["model_frac_of_bf16_peak"],
            "roofline": head["roofline"], "kernels": head["kernels"], "cpu_baseline": cpu, "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"] + sum(r["gpu_launches"] for r in subs.values() if r),
            "clocks": head["clocks"], "host_binding": host_binding,
        }
        if "e2e_uint8_inputs" in head:
            line["e2e_uint8_inputs"] = head["e2e_uint8_inputs"]
        for k in ("allreduce_exposed_ms", "ms_per_step_without_allreduce", "allreduce_bytes_per_step", "reserve_sms"):
            if k in head:
                line[k] = head[k]
        for name, r in subs.items():
            if r is None:
                continue
            r = dict(r)
            r["scaling"] = "strong" if name == "encoder_strong" else "weak"
            r["dtype"], r["data"] = "bf16", "synthetic"
            line[name] = r
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="default", choices=["default"] + sorted(WORKLOADS),
                    help="default = encoder_large headline + pretrain_large (+ encoder_strong when N > 1)")
    ap.add_argument("--skip-pretrain", action="store_true", help="default workload: encoder line only")
    ap.add_argument("--skip-strong", action="store_true", help="default workload: no strong-scaling encoder point")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="training steps: optim.FusedAdamW (csrc/optim.cu) or torch.optim.AdamW(fused=True)")
    ap.add_argument("--mask-sampler", default="device", choices=["device", "reference"],
                    help="pretraining: fused on-device mask sampling kernel or the reference's op sequence")
    ap.add_argument("--bucket-mb", type=float, default=64.0, help="gradient all-reduce bucket size")
    ap.add_argument("--linear-probe", action="store_true",
                    help="cls_large: only head.* trains (fm_cls_config.py:111-124)")
    ap.add_argument("--label-smoothing", type=float, default=0.0,
                    help="cls_large: label-smoothing cross entropy (run_cls_tuning.py:438-441)")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="SMs kept free of persistent compute CTAs while gradient buckets are in flight "
                         "(-1 = auto: 16 when N > 1; measured at N = 2: exposed all-reduce 3.7 / 1.9 / 1.4 ms with 0 / 8 / 16)")
    ap.add_argument("--nccl-max-ctas", type=int, default=-1,
                    help="NCCL_MAX_CTAS for the gradient all-reduce (-1 = auto: same as --reserve-sms, 0 = leave)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every kernel eagerly")
    # e2e leg: [ramp, everything else, ramp] images; 18 / 220 images x 513 tokens are whole numbers of 256-row
    # tiles per wave, and three chunks keep the per-kernel ramp-up cost (195 launches per chunk) small
    ap.add_argument("--e2e-chunk", type=int, default=0, help="images per middle chunk of the e2e leg (0 = one)")
    ap.add_argument("--e2e-ramp", type=int, default=18, help="images in the first and last (short) chunk")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "default":
            args.workload = "encoder_large"
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = max(world, args.gpus)
    if args.reserve_sms < 0:
        args.reserve_sms = 16 if n > 1 else 0
    if args.nccl_max_ctas < 0:
        args.nccl_max_ctas = args.reserve_sms
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", __file__] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
