#!/usr/bin/env python
"""bench.py — images/s of the reduced-ViT forward with the tokred reduction kernels (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl tokred|reference] [--workload NAME] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one forward pass of the hot path's host model over one synthetic batch (random-init weights, N(0,1)
224x224 images): the backbone runs on PyTorch/cuBLAS exactly like the reference's, the reduction operators at
blocks 3/6/9 are the hand-written sm_100a kernels of libtokred_sm100a.so.  Default workload = BASELINE.json
configs[1] (DeiT-S ToMe, reduction_loc 3 6 9, batch 256 per GPU, bf16 autocast).  Multi-GPU = pure batch sharding
(weak scaling: the per-GPU batch is fixed), with one NCCL all_gather of the logits and of the last stage's
assignment map per step — the only exchange the path has.

One JSON line on rank 0:
  value      images/s, all ranks, inputs resident in HBM, CUDA events, max over ranks
  e2e        same through the public API with pinned-host inputs: H2D of the batch + D2H of the logits per step
  roofline   dominant tokred kernel (largest stage): algorithmic bytes / measured duration vs MEASURED_PEAKS hbm_gbs
  kernels    every tokred launch of one step: avg us, algorithmic GB/s, fraction of peak
  cpu_baseline  the oracle port of the reference model on the host cores, bounded sample
--impl reference: the oracle port (CPU restatement pinned to the reference, oracle/model.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (factory method, size, keep_rate, per-GPU batch, amp)   — BASELINE.json configs
    "topk_small_kr0.7_b64": ("topk", "small", 0.7, 64, False),
    "tome_small_kr0.7_b256_bf16": ("tome", "small", 0.7, 256, True),
    "evit_base_kr0.5_b128": ("evit", "base", 0.5, 128, True),
    "dyvit_base_kr0.5_b128": ("dyvit", "base", 0.5, 128, True),
    "dpcknn_small_kr0.25_b256": ("dpcknn", "small", 0.25, 256, True),
    "kmedoids_small_kr0.25_b256": ("kmedoids", "small", 0.25, 256, True),
    "ats_base_kr0.9_b128": ("ats", "base", 0.9, 128, True),
    "sinkhorn_base_kr0.9_b128": ("sinkhorn", "base", 0.9, 128, True),
    "patchmerger_base_kr0.9_b128": ("patchmerger", "base", 0.9, 128, True),
    "sit_base_kr0.9_b128": ("sit", "base", 0.9, 128, True),
}
DEFAULT_WORKLOAD = "tome_small_kr0.7_b256_bf16"
DIMS = {"tiny": (192, 3), "small": (384, 6), "base": (768, 12)}


def model_args(kr):
    return Namespace(keep_rate=[kr], reduction_loc=[3, 6, 9], distillation_type="none", k_neighbors=5, cluster_iters=3,
                     sinkhorn_eps=1.0, equal_weight=False, dyvit_distill=False)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled in the background DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ algorithmic bytes
_ABI_ARGS = None


def abi_arg_names():
    """{entry point: [parameter names]} parsed from include/tokred.h, so that the byte formulas below address the
    ctypes argument tuples by NAME (a positional table silently broke when x_batch_stride joined six signatures)."""
    global _ABI_ARGS
    if _ABI_ARGS is None:
        import re
        text = open(os.path.join(ROOT, "include", "tokred.h")).read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        _ABI_ARGS = {}
        for m in re.finditer(r"TOKRED_API\s+[\w\s\*]+?\b(tokred_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
            params = [q.strip() for q in m.group(2).split(",") if q.strip() and q.strip() != "void"]
            _ABI_ARGS[m.group(1)] = [re.findall(r"\w+", q)[-1] for q in params]
    return _ABI_ARGS


def algorithmic_bytes(name: str, a: tuple) -> float:
    """Bytes one launch must move (inputs read once + outputs written once), from the ctypes argument tuple.
    Formulas: SURVEY.md §8(d) / DESIGN.md §3."""
    names = abi_arg_names().get(name)
    if not names or len(names) != len(a):
        return 0.0
    g = dict(zip(names, a))

    def esz(dt):
        return 2 if dt == 1 else 4
    if name == "tokred_tome_match":
        n, r = g["N"], g["r"]
        na = (n + 1) // 2
        re_ = min(r, (n - 1) // 2)
        return g["B"] * (n * g["D"] * esz(g["metric_dtype"]) + 8 * (na - re_) + 16 * re_)
    if name == "tokred_tome_merge":
        n, c, r, e = g["N"], g["C"], g["r"], esz(g["x_dtype"])
        na = (n + 1) // 2
        total = n * c * e + (n - r) * c * e + (n - r) * e + 8 * (na - r) + 16 * r
        if g["size"]:
            total += n * e
        if g["reduced_cluster_idx"]:
            total += 4 * (n - 1)
        return g["B"] * total
    if name == "tokred_topk_gather":
        n, c, k = g["N"], g["C"], g["k"]
        sc = (n - 1) * esz(g["score_dtype"]) if g["scores"] else g["H"] * (n - 1) * esz(g["attn_dtype"])
        return g["B"] * (sc + 2 * (k + 1) * c * esz(g["x_dtype"]) + 8 * k)
    if name == "tokred_evit_select_fuse":
        n, c, k = g["N"], g["C"], g["k"]
        sc = (n - 1) * esz(g["score_dtype"]) if g["scores"] else g["H"] * (n - 1) * esz(g["attn_dtype"])
        return g["B"] * (sc + n * c * esz(g["x_dtype"]) + (k + 2) * c * esz(g["x_dtype"]) + 8 * (k + 1) + 8 * (n - 1 - k))
    if name == "tokred_dpcknn_cluster":
        p, c, k = g["P"], g["C"], g["K"]
        return g["B"] * (p * c * 4 + p * 4 + 8 * p + 8 * k)
    if name == "tokred_dpcknn_merge":
        p, c, k, t = g["P"], g["C"], g["K"], g["T"]
        return g["B"] * (p * c * 4 + p * 4 + 8 * p + k * c * 4 + t * (8 + 4) * 2)
    if name == "tokred_kmedoids_fit":
        p, c, k = g["P"], g["C"], g["K"]
        return g["B"] * (p * c * 4 + p * 4 + k * c * 4 + 8 * k + 8 * p)
    if name == "tokred_attn_colsum":
        n = g["N"]
        return g["B"] * (g["H"] * n * n * esz(g["attn_dtype"]) + 4 * (n - 1))
    if name in ("tokred_sinkhorn_merge", "tokred_patchmerger"):
        p, c, k = g["P"], g["C"], g["K"]
        return g["B"] * (p * c * esz(g["x_dtype"]) + k * c * esz(g["out_dtype"]) + k * p * 4) + k * c * 4
    if name == "tokred_sit_merge":
        p, c, k = g["P"], g["C"], g["K"]
        return g["B"] * (p * c * esz(g["x_dtype"]) + p * k * esz(g["logits_dtype"]) + k * c * esz(g["out_dtype"]) + k * p * 4)
    if name == "tokred_ats_sample":
        n, h = g["N"], g["H"]
        return g["B"] * (h * (n - 1) * 4 + h * n * g["Dh"] * esz(g["v_dtype"]) + n + 9 * (g["n_steps"] + 1))
    if name == "tokred_gather_rows":
        m = g["M"]
        return g["B"] * (2 * g["G"] * m * g["W"] * esz(g["dtype"]) + 8 * m)
    if name == "tokred_dyvit_pool_concat":
        p, c = g["P"], g["C"]
        return g["B"] * (p * c * esz(g["h_dtype"]) + p * 4 + p * c * esz(g["out_dtype"]))
    return 0.0


def summarise_timeline(timeline, steps, peak_gbs):
    """group launches by (name, shape) -> avg duration, algorithmic GB/s, fraction of the HBM peak."""
    groups = {}
    for name, args, e0, e1 in timeline:
        key = (name, tuple(x for x in args if isinstance(x, int) and not isinstance(x, bool) and abs(x) < (1 << 24)))
        g = groups.setdefault(key, {"name": name, "ms": [], "bytes": algorithmic_bytes(name, args)})
        g["ms"].append(e0.elapsed_time(e1))
    rows = []
    for (name, _), g in groups.items():
        avg_ms = sum(g["ms"]) / len(g["ms"])
        gbs = g["bytes"] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        rows.append({"kernel": name.replace("tokred_", ""), "launches_per_step": len(g["ms"]) / max(steps, 1),
                     "avg_us": round(avg_ms * 1e3, 2), "alg_mb": round(g["bytes"] / 1e6, 3), "alg_gbs": round(gbs, 1),
                     "frac_hbm": round(gbs / peak_gbs, 4)})
    rows.sort(key=lambda r: -r["alg_mb"])
    return rows


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(method, size, kr, amp, sample_batch, steps, warmup):
    """The oracle port of the reference model (oracle/model.py, pinned to the unmodified reference) on the host
    cores, fp32, all threads.  Returns (images/s, ms/step, cores)."""
    from oracle import model as OM
    from tokenreduction_b200 import create_model
    import contextlib
    import io
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        sd = create_model(f"{method}_{size}_patch16_224", num_classes=1000, args=model_args(kr)).state_dict()
    sd = {k: v.detach().clone() for k, v in sd.items()}
    cfg = OM.cfg_for(size, keep_rate=[kr])
    x = torch.randn(sample_batch, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    for _ in range(warmup):
        OM.forward(method, sd, x, cfg)
    t0 = time.perf_counter()
    for _ in range(steps):
        OM.forward(method, sd, x, cfg)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_batch / dt, dt * 1e3, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    method, size, kr, batch, amp = WORKLOADS[a.workload]
    sample = min(batch, a.cpu_batch)
    ips, ms, cores = cpu_reference_run(method, size, kr, amp, sample, a.steps, max(a.warmup, 1))
    line = {
        "impl": "reference", "metric": "images_per_s", "value": round(ips, 2), "unit": "images/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": a.workload, "per_gpu_batch": batch, "reduction_loc": [3, 6, 9], "keep_rate": kr,
                   "note": "oracle port of the reference model on host cores; each step = one forward over the sample"},
        "cpu_baseline": {"value": round(ips, 2), "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images/step x {a.steps} steps, fp32, torch CPU threads={cores}"},
        "e2e": {"value": round(ips, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ tokred arm
def run_tokred(a):
    import torch.distributed as dist
    from tokenreduction_b200 import _lib, create_model
    import contextlib
    import io

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the tokred kernels have no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    method, size, kr, batch, amp = WORKLOADS[a.workload]
    if a.batch:
        batch = a.batch
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(f"{method}_{size}_patch16_224", pretrained=False, num_classes=1000, drop_rate=0.0,
                             drop_path_rate=0.0, drop_block_rate=None, img_size=224, args=model_args(kr))
    model = model.eval().to(dev)
    gen = torch.Generator().manual_seed(1 + rank)
    host_images = torch.randn(batch, 3, 224, 224, generator=gen).pin_memory()
    images = host_images.to(dev)
    host_logits = torch.empty(batch, 1000, dtype=torch.float32).pin_memory()
    gathered = [torch.empty(batch, 1000, device=dev) for _ in range(world)] if world > 1 else None

    def forward(x):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            y = model(x).float()
        if world > 1:                      # the path's only exchange: gather the logits of every shard
            dist.all_gather(gathered, y)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        forward(images)
    barrier()

    peaks, peak_kind = measured_peaks()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---- timed region 1: inputs resident in HBM (value); every tokred launch bracketed by events on its stream
    _lib.TIMELINE = []
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        forward(images)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    timeline, _lib.TIMELINE = _lib.TIMELINE, None
    kernels = summarise_timeline(timeline, a.steps, peaks["hbm_gbs"])

    # ---- timed region 2: end to end through the public API (pinned host -> device, forward, logits -> host).
    # Every step copies ITS batch from pinned host memory and reads ITS logits back; the copy of step i+1 runs on a
    # second stream into the other of two device buffers while step i computes (what a serving loop does).
    cur = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(images), torch.empty_like(images)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]      # H2D into buffer i finished
    freed = [None, None]                                  # forward that read buffer i finished

    def stage(i, after=None):
        with torch.cuda.stream(copy_stream):
            if after is not None:
                copy_stream.wait_event(after)
            if freed[i % 2] is not None:
                copy_stream.wait_event(freed[i % 2])
            bufs[i % 2].copy_(host_images, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_steps(n, start_event=None):
        stage(0, start_event)
        for i in range(n):
            if i + 1 < n:
                stage(i + 1)
            cur.wait_event(ready[i % 2])
            y = forward(bufs[i % 2])
            freed[i % 2] = torch.cuda.Event()
            freed[i % 2].record(cur)
            host_logits.copy_(y, non_blocking=True)
            cur.synchronize()                              # the caller reads the logits every step

    e2e_steps(2)
    barrier()
    freed = [None, None]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(a.steps, e0)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clock_rec = clocks.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        total = batch * world
        top = kernels[0] if kernels else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if top and os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(a.workload, {}).get(top["kernel"])
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            sample = min(batch, a.cpu_batch)
            ips, cms, cores = cpu_reference_run(method, size, kr, amp, sample, a.cpu_steps, 1)
            cpu = {"value": round(ips, 2), "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": f"{sample} images/step x {a.cpu_steps} steps of the same model (oracle port, fp32), "
                             f"{cms:.0f} ms/step"}
        line = {
            "metric": "images_per_s", "value": round(total * a.steps / (ms * 1e-3), 1), "unit": "images/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if amp else "f32",
            "data": "synthetic",
            "config": {"workload": a.workload, "model": f"{method}_{size}_patch16_224", "per_gpu_batch": batch,
                       "global_batch": total, "keep_rate": kr, "reduction_loc": [3, 6, 9], "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (batch of fp32 images = %.0f MB)" % (batch * 3 * 224 * 224 * 4 / 1e6),
                       "timing": "CUDA events, max over ranks",
                       "e2e_pipeline": "per step: H2D of the batch (pinned, copy stream, 2 device buffers; overlaps the "
                                       "previous step's forward) + forward + D2H of the logits + stream sync"},
            "e2e": {"value": round(total * a.steps / (ms_e2e * 1e-3), 1), "unit": "images/s",
                    "h2d_bytes_per_step": host_images.numel() * 4 * world, "d2h_bytes_per_step": host_logits.numel() * 4 * world},
            "gpu_launches": int(launches),
            "clocks": clock_rec,
            "roofline": None if not top else {
                "kernel": top["kernel"], "bound": "hbm", "achieved": top["alg_gbs"], "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": top["frac_hbm"], "traffic": traffic, "peak_source": peak_kind,
                "avg_us": top["avg_us"], "alg_mb_per_launch": top["alg_mb"]},
            "kernels": kernels,
            "tokred_share_of_step": round(sum(k["avg_us"] * k["launches_per_step"] for k in kernels) / (ms / a.steps * 1e3), 4),
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tokred", choices=["tokred", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=32, help="images per step of the CPU baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called plainly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_tokred(a)


if __name__ == "__main__":
    main()
