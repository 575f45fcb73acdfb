#!/usr/bin/env python
"""bench.py — images/s of the reduced-ViT forward with the tokred reduction kernels (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl tokred|reference] [--workload NAME] [--batch B]
                    [--headline-only]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one forward pass of the hot path's host model over one synthetic batch (random-init weights, N(0,1)
224x224 images): the backbone runs on PyTorch/cuBLAS exactly like the reference's, the reduction operators at
blocks 3/6/9 are the hand-written sm_100a kernels of libtokred_sm100a.so.  Headline workload = BASELINE.json
configs[1] (DeiT-S ToMe, reduction_loc 3 6 9, batch 256 per GPU, bf16 autocast), weak scaling (per-GPU batch fixed).
Multi-GPU = pure batch sharding with one NCCL exchange per step: all_gather of the logits AND of every stage's
kept-token / assignment indices (north_star; the reference's consumer is validate.py:199-229).

One JSON line on rank 0:
  value      images/s of the headline, all ranks, inputs resident in HBM, CUDA events, max over ranks
  e2e        same through the public API with pinned-host inputs: H2D of the batch + D2H of the logits per step
  roofline   dominant tokred kernel (largest stage): algorithmic bytes / measured duration vs MEASURED_PEAKS hbm_gbs
  kernels    every tokred launch of one step: avg us, algorithmic GB/s, fraction of peak
  cpu_baseline  the unmodified reference (baseline/_ref, else the oracle port) on the host cores, bounded sample
  workloads  every other BASELINE.json config measured the same way in the same run: config 1 (Top-K S B=64), config 3
             (EViT / DynamicViT B, global batch 1024 STRONG-scaled over the ranks), config 4 (DPC-KNN / K-Medoids S
             B=256), config 5 (ATS / Sinkhorn / PatchMerger B, per-GPU batch sweep), SiT
--impl reference: the reference's own CPU implementation on the host cores (kind "reference" when baseline/_ref is
staged, "port" = oracle/model.py otherwise), at the headline's per-GPU batch.
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


def extra_workloads(world):
    """(label, BASELINE config, workload key, per-GPU batch, scaling) measured after the headline in the same run."""
    out = [("topk_small_kr0.7_b64", 1, "topk_small_kr0.7_b64", 64, "weak")]
    for w in ("evit_base_kr0.5_b128", "dyvit_base_kr0.5_b128"):         # config 3: batch 1024 SHARDED over the GPUs
        out.append((w.replace("_b128", f"_global1024_dp{world}"), 3, w, max(1024 // world, 1), "strong"))
    for w in ("dpcknn_small_kr0.25_b256", "kmedoids_small_kr0.25_b256"):
        out.append((w, 4, w, 256, "weak"))
    for w in ("ats_base_kr0.9_b128", "sinkhorn_base_kr0.9_b128", "patchmerger_base_kr0.9_b128"):   # config 5: sweep
        for bsz in (128, 512, 1024):
            out.append((w.replace("_b128", f"_b{bsz}"), 5, w, bsz, "weak"))
    out.append(("sit_base_kr0.9_b128", None, "sit_base_kr0.9_b128", 128, "weak"))
    return out


def model_args(kr, **kw):
    return Namespace(keep_rate=[kr], reduction_loc=[3, 6, 9], distillation_type="none", k_neighbors=5, cluster_iters=3,
                     sinkhorn_eps=1.0, equal_weight=False, dyvit_distill=False, **kw)


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
        return g["B"] * (n * g["D"] * max(1, g["heads"]) * esz(g["metric_dtype"]) + 8 * (na - re_) + 16 * re_)
    if name == "tokred_tome_merge":
        n, c, r, e = g["N"], g["C"], g["r"], esz(g["x_dtype"])
        na = (n + 1) // 2
        total = n * c * e + (n - r) * c * e + (n - r) * e + 8 * (na - r) + 16 * r
        if g["size"]:
            total += n * e
        if g["reduced_cluster_idx"]:
            total += 4 * (n - 1)
        return g["B"] * total
    if name == "tokred_tome_merge_ln":
        n, c, r = g["N"], g["C"], g["r"]
        na = (n + 1) // 2
        total = n * c * 4 + (n - r) * c * (4 + 2) + (n - r) * 4 + 8 * (na - r) + 16 * r + 8 * c   # x, x_out + y, sizes, lists
        if g["branch"]:
            total += n * c * 2
        if g["size"]:
            total += n * 4
        if g["reduced_cluster_idx"]:
            total += 4 * (n - 1)
        return g["B"] * total
    if name == "tokred_topk_gather":
        n, c, k = g["N"], g["C"], g["k"]
        sc = (n - 1) * esz(g["score_dtype"]) if g["scores"] else g["H"] * (n - 1) * esz(g["attn_dtype"])
        return g["B"] * (sc + 2 * (k + 1) * c * esz(g["x_dtype"]) + 8 * k)
    if name == "tokred_topk_gather_add":
        n, c, k = g["N"], g["C"], g["k"]
        return g["B"] * ((n - 1) * esz(g["score_dtype"]) + (k + 1) * c * (4 + 2 + 4) + 8 * k)
    if name == "tokred_evit_select_fuse_add":
        n, c, k = g["N"], g["C"], g["k"]
        return g["B"] * ((n - 1) * esz(g["score_dtype"]) + n * c * (4 + 2) + (k + 2) * c * 4 + 8 * (k + 1) + 8 * (n - 1 - k))
    if name == "tokred_residual_add":
        return g["n"] * (4 + 2 + 4)
    if name == "tokred_patchify":
        return g["B"] * g["Cin"] * g["H"] * g["W"] * (4 + 2)
    if name == "tokred_embed_layernorm":
        n = g["T"] + g["P"]
        return g["B"] * (g["P"] * g["C"] * 2 + n * g["C"] * (4 + 2)) + (n + g["T"] + 2) * g["C"] * 4
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
    if name == "tokred_add_layernorm":
        per = g["C"] * (4 + 2)                                   # x read, y written
        if g["branch"]:
            per += g["C"] * (esz(g["branch_dtype"]) + 4)          # branch read, new residual row written
        return g["rows"] * per + 8 * g["C"]
    if name == "tokred_attention":
        n, hh, m = g["N"], g["H"], (g["M"] if g["q_ids"] else g["N"])
        c = hh * g["head_dim"]
        per = m * c * 2 + n * c * 2                      # q rows, k
        if g["out"]:
            per += n * c * 2 + m * c * 2                 # v, out
        per += (4 * n if g["key_bias"] else 0) + (n if g["mask"] else 0) + (8 * m if g["q_ids"] else 0)
        per += (4 * hh * n if g["cls_row"] else 0) + (4 * hh * n if g["colsum"] else 0)
        return g["B"] * per
    if name == "tokred_dyvit_pool_concat":
        p, c = g["P"], g["C"]
        return g["B"] * (p * c * esz(g["h_dtype"]) + p * 4 + p * c * esz(g["out_dtype"]))
    return 0.0


def dominant_kernel(rows):
    """the roofline object describes the DOMINANT tokred kernel: the kernel family with the largest share of the step
    (sum of avg_us x launches over its launch groups), represented by its largest launch group (first stage)."""
    if not rows:
        return None
    share = {}
    for r in rows:
        share[r["kernel"]] = share.get(r["kernel"], 0.0) + r["avg_us"] * r["launches_per_step"]
    fam = max(share, key=share.get)
    return max((r for r in rows if r["kernel"] == fam), key=lambda r: r["alg_mb"])


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
    """The reference model on the host cores, fp32, all threads.  kind "reference": the UNMODIFIED reference classes
    from baseline/_ref (staged by __graft_entry__.build(); imported through the timm-0.4.12 API shim, the only missing
    dependency) built by the reference's own factory; kind "port": the oracle port (oracle/model.py, pinned to the
    reference) when no staged copy exists.  Returns (images/s, ms/step, cores, kind)."""
    import contextlib
    import io
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.randn(sample_batch, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    from oracle import timm_shim
    if timm_shim.reference_available():
        kind = "reference"
        models_act = timm_shim.import_reference()
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = timm_shim.create_model(f"{method}_{size}_patch16_224", pretrained=False, num_classes=1000, drop_rate=0.0,
                                           drop_path_rate=0.0, drop_block_rate=None, img_size=224, args=model_args(kr)).eval()
        del models_act

        def run():
            with torch.no_grad():
                return model(x)
    else:
        kind = "port"
        from oracle import model as OM
        from tokenreduction_b200 import create_model
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            sd = create_model(f"{method}_{size}_patch16_224", num_classes=1000, args=model_args(kr)).state_dict()
        sd = {k: v.detach().clone() for k, v in sd.items()}
        cfg = OM.cfg_for(size, keep_rate=[kr])

        def run():
            return OM.forward(method, sd, x, cfg)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_batch / dt, dt * 1e3, torch.get_num_threads(), kind


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    method, size, kr, batch, amp = WORKLOADS[a.workload]
    if a.batch:
        batch = a.batch
    sample = min(batch, a.cpu_batch) if a.cpu_batch else batch
    ips, ms, cores, kind = cpu_reference_run(method, size, kr, amp, sample, a.steps, max(min(a.warmup, 2), 1))
    line = {
        "impl": "reference", "metric": "images_per_s", "value": round(ips, 2), "unit": "images/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": a.workload, "model": f"{method}_{size}_patch16_224", "per_gpu_batch": batch,
                   "keep_rate": kr, "reduction_loc": [3, 6, 9],
                   "note": ("unmodified reference model (baseline/_ref)" if kind == "reference" else "oracle port of the reference model")
                           + f" on the host cores, fp32 (the GPU arm runs bf16 autocast), {sample} images per step"},
        "cpu_baseline": {"value": round(ips, 2), "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} images/step x {a.steps} steps, fp32, torch CPU threads={cores}"},
        "e2e": {"value": round(ips, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ tokred arm
DECISION_KEYS = ("Kept_Tokens", "Assignment_Maps")


def pack_decisions(viz, logits=None):
    """the logits (fp32 bit patterns) and every stage's kept-token / assignment indices of this rank's shard as ONE int32
    tensor [B, total] (device)."""
    cols = [] if logits is None else [logits.contiguous().view(torch.int32)]
    for key in DECISION_KEYS:
        for i in sorted(viz.get(key, {})):
            t = viz[key][i]
            cols.append(t.reshape(t.shape[0], -1).to(torch.int32))
    return torch.cat(cols, dim=1) if cols else None


class Runner:
    """one workload on this rank: model, synthetic shard, forward (+ the NCCL exchange when world > 1)."""

    def __init__(self, wl_key, batch, dev, rank, world):
        import contextlib
        import io
        from tokenreduction_b200 import create_model
        import torch.distributed as dist
        self.dist, self.world, self.dev = dist, world, dev
        method, size, kr, _, amp = WORKLOADS[wl_key]
        self.method, self.size, self.kr, self.amp, self.batch = method, size, kr, amp, batch
        torch.backends.cudnn.benchmark = True      # the reference's own evaluation setting (validate.py:32,63)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            # viz_mode + tokred_device_viz: the per-stage decisions come back as DEVICE tensors (no copy, no sync)
            model = create_model(f"{method}_{size}_patch16_224", pretrained=False, num_classes=1000, drop_rate=0.0,
                                 drop_path_rate=0.0, drop_block_rate=None, img_size=224,
                                 args=model_args(kr, viz_mode=world > 1, tokred_device_viz=True))
        self.model = model.eval().to(dev)
        self.model.viz_features = False        # only the kept / assignment indices are gathered across ranks, not feature maps
        gen = torch.Generator().manual_seed(1 + rank)
        self.host_images = torch.randn(batch, 3, 224, 224, generator=gen).pin_memory()
        self.images = self.host_images.to(dev)
        self.host_logits = torch.empty(batch, 1000, dtype=torch.float32).pin_memory()
        self.g_logits = torch.empty(world * batch, 1000, device=dev) if world > 1 else None
        self.g_dec = None
        self.dec_cols = 0
        self.graphs = {}

    def graphed(self, buf):
        """one CUDA graph of the forward reading `buf` in place (tokenreduction_b200.graph: a step = one launch)."""
        from tokenreduction_b200.graph import GraphedForward
        g = self.graphs.get(buf.data_ptr())
        if g is None:
            g = GraphedForward(self._step, buf, torch.bfloat16 if self.amp else None, static_input=buf)
            self.graphs[buf.data_ptr()] = g
        return g

    def _step(self, x):
        """the forward as it is captured into the CUDA graph.  world > 1: the model runs in viz mode and the shard's logits
        (fp32 bit patterns) and kept / assignment indices are packed into ONE int32 tensor [B, 1000 + index columns] inside
        the same graph, so the only work outside it is the all_gather."""
        out = self.model(x)
        if self.world == 1:
            return out.float()
        y, viz = out
        y = y.float()
        return y, pack_decisions(viz, y)

    def forward(self, x, graph=False):
        if graph:
            out = self.graphed(x)(x)
        else:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                out = self._step(x)
        if self.world == 1:
            return out
        y, dec = out
        # the path's only exchange: one all_gather per step of the packed tensor (KBs-MBs)
        if self.g_dec is None or self.g_dec.shape[1] != dec.shape[1]:
            self.g_dec = torch.empty(self.world * dec.shape[0], dec.shape[1], dtype=torch.int32, device=self.dev)
            self.dec_cols = dec.shape[1] - y.shape[1]
        self.dist.all_gather_into_tensor(self.g_dec, dec)
        self.g_logits = self.g_dec[:, :y.shape[1]].view(torch.float32)     # every rank holds all logits and all indices
        return y

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def measure(self, steps, warmup, timeline=False, e2e=True, graph=True):
        """-> dict(ms, ms_e2e, launches, kernels timeline).  Timed region 1: inputs resident in HBM.  Timed region 2: end
        to end through the public API -- every step copies ITS batch from pinned host memory and reads ITS logits back;
        the copy of step i+1 runs on a second stream into the other of two device buffers while step i computes."""
        from tokenreduction_b200 import _lib
        for _ in range(max(warmup, 3)):
            self.forward(self.images)
        self.barrier()
        # per-kernel timeline: a separate, untimed eager pass with an event pair around every tokred launch (the event
        # records cost host time, so they stay out of the timed region); also counts the launches of one step
        res = {"timeline": None, "ms_e2e": None}
        tl_steps = min(steps, 3)
        if timeline:
            _lib.TIMELINE = []
        launches0 = _lib.launch_count()
        for _ in range(tl_steps):
            self.forward(self.images)
        self.barrier()
        res["launches"] = (_lib.launch_count() - launches0) * steps // tl_steps
        res["timeline_steps"] = tl_steps
        if timeline:
            res["timeline"], _lib.TIMELINE = _lib.TIMELINE, None
        if graph:
            for _ in range(2):
                self.forward(self.images, True)      # capture + first replays
            self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(steps):
            self.forward(self.images, graph)
        ev1.record()
        self.barrier()
        res["ms"] = ev0.elapsed_time(ev1)
        if not e2e:
            return res
        cur = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        bufs = [torch.empty_like(self.images), torch.empty_like(self.images)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]      # H2D into buffer i finished
        freed = [None, None]                                  # forward that read buffer i finished

        def stage(i, after=None):
            with torch.cuda.stream(copy_stream):
                if after is not None:
                    copy_stream.wait_event(after)
                if freed[i % 2] is not None:
                    copy_stream.wait_event(freed[i % 2])
                bufs[i % 2].copy_(self.host_images, non_blocking=True)
                ready[i % 2].record(copy_stream)

        host_out = [self.host_logits, torch.empty_like(self.host_logits).pin_memory()]
        landed = [None, None]                                 # D2H of step i's logits finished

        def e2e_steps(n, start_event=None):
            stage(0, start_event)
            for i in range(n):
                if i + 1 < n:
                    stage(i + 1)
                cur.wait_event(ready[i % 2])
                y = self.forward(bufs[i % 2], graph)
                freed[i % 2] = torch.cuda.Event()
                freed[i % 2].record(cur)
                host_out[i % 2].copy_(y, non_blocking=True)
                landed[i % 2] = torch.cuda.Event()
                landed[i % 2].record(cur)
                # the caller reads every step's logits: step i-1's while step i is already enqueued (the device never waits
                # for the host between two steps), the last step's after the loop
                if i >= 1:
                    landed[(i - 1) % 2].synchronize()
            landed[(n - 1) % 2].synchronize()

        e2e_steps(2)
        self.barrier()
        freed[0] = freed[1] = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_steps(steps, e0)
        e1.record()
        self.barrier()
        res["ms_e2e"] = e0.elapsed_time(e1)
        return res


def bind_to_gpu_numa_node(dev):
    """Multi-GPU end-to-end path: every rank streams its 154 MB batch from pinned host memory each step.  Pinned pages are
    placed on the NUMA node of the thread that first touches them, so the rank is bound to the CPUs NVML reports as local to
    its GPU BEFORE it allocates the pinned batch -- otherwise about half of the eight copies cross the socket interconnect.
    Returns the CPU count bound to, or None (no NVML / TOKRED_BENCH_NO_AFFINITY=1 / any error: the default placement stays)."""
    if os.environ.get("TOKRED_BENCH_NO_AFFINITY") == "1" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1} & allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def run_tokred(a):
    import torch.distributed as dist
    from tokenreduction_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the tokred kernels have no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(dev) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    def max_over_ranks(*vals):
        if world == 1:
            return vals
        t = torch.tensor([float("nan") if v is None else v for v in vals], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in t)

    method, size, kr, batch, amp = WORKLOADS[a.workload]
    if a.batch:
        batch = a.batch
    peaks, peak_kind = measured_peaks()
    run = Runner(a.workload, batch, dev, rank, world)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    res = run.measure(a.steps, a.warmup, timeline=True, graph=not a.no_graph)
    clock_rec = clocks.stop() if rank == 0 else None
    ms, ms_e2e = max_over_ranks(res["ms"], res["ms_e2e"])
    kernels = summarise_timeline(res["timeline"], res["timeline_steps"], peaks["hbm_gbs"])
    launches = res["launches"]
    dec_cols = run.dec_cols
    h2d = run.host_images.numel() * 4 * world
    d2h = run.host_logits.numel() * 4 * world
    del run
    torch.cuda.empty_cache()

    # ---- every other BASELINE config, same method, shorter (they are reported in "workloads", not as the headline)
    extras = []
    if not a.headline_only:
        xsteps, xwarm = max(3, min(a.steps, 6)), 3
        for label, cfg_no, key, bsz, scaling in extra_workloads(world):
            try:
                r = Runner(key, bsz, dev, rank, world)
                rr = r.measure(xsteps, xwarm, timeline=True, graph=not a.no_graph)
                xms, xms_e2e = max_over_ranks(rr["ms"], rr["ms_e2e"])
                ks = summarise_timeline(rr["timeline"], rr["timeline_steps"], peaks["hbm_gbs"])
                top = dominant_kernel(ks)
                m_, s_, kr_, _, amp_ = WORKLOADS[key]
                extras.append({
                    "workload": label, "baseline_config": cfg_no, "model": f"{m_}_{s_}_patch16_224", "keep_rate": kr_,
                    "dtype": "bf16" if amp_ else "f32", "per_gpu_batch": bsz, "global_batch": bsz * world, "scaling": scaling,
                    "value": round(bsz * world * xsteps / (xms * 1e-3), 1), "unit": "images/s",
                    "ms_per_step": round(xms / xsteps, 3), "e2e": round(bsz * world * xsteps / (xms_e2e * 1e-3), 1),
                    "gpu_launches_per_step": rr["launches"] / xsteps,
                    "top_kernel": None if not top else {"kernel": top["kernel"], "avg_us": top["avg_us"], "alg_mb": top["alg_mb"],
                                                        "alg_gbs": top["alg_gbs"], "frac_hbm": top["frac_hbm"]},
                    "kernels": [{k: v for k, v in kk.items() if k in ("kernel", "launches_per_step", "avg_us", "frac_hbm")} for kk in ks],
                    "tokred_share_of_step": round(sum(k["avg_us"] * k["launches_per_step"] for k in ks) / (xms / xsteps * 1e3), 4),
                })
                del r
            except Exception as e:      # one failing workload must not cost the headline line
                extras.append({"workload": label, "baseline_config": cfg_no, "failed": f"{type(e).__name__}: {e}"[:300]})
            torch.cuda.empty_cache()

    if rank == 0:
        total = batch * world
        top = dominant_kernel(kernels)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if top and os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(a.workload, {}).get(top["kernel"])
            if isinstance(traffic, dict):      # {"bytes": ..., "grid": ..., "capture": ...} written by tools/ncu_traffic.py
                traffic = traffic.get("bytes")
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            sample = min(batch, a.cpu_batch) if a.cpu_batch else batch
            ips, cms, cores, kind = cpu_reference_run(method, size, kr, amp, sample, a.cpu_steps, 1)
            cpu = {"value": round(ips, 2), "unit": "images/s", "cores": cores, "kind": kind,
                   "sample": f"{sample} images/step x {a.cpu_steps} steps of the same model "
                             f"({'unmodified reference, baseline/_ref' if kind == 'reference' else 'oracle port'}, fp32), {cms:.0f} ms/step"}
        line = {
            "metric": "images_per_s", "value": round(total * a.steps / (ms * 1e-3), 1), "unit": "images/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if amp else "f32",
            "data": "synthetic",
            "config": {"workload": a.workload, "model": f"{method}_{size}_patch16_224", "per_gpu_batch": batch,
                       "global_batch": total, "keep_rate": kr, "reduction_loc": [3, 6, 9], "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (batch of fp32 images = %.0f MB)" % (batch * 3 * 224 * 224 * 4 / 1e6),
                       "timing": "CUDA events, max over ranks",
                       "execution": "eager (one Python-issued launch per kernel)" if a.no_graph else
                                    "one CUDA graph replay per step (tokenreduction_b200.graph.GraphedForward); the per-kernel "
                                    "timeline is a separate untimed eager pass",
                       "host_affinity": (f"rank bound to the {affinity} CPUs NVML reports local to its GPU before the pinned batch is "
                                         "allocated (first-touch NUMA placement)") if affinity else "default",
                       "exchange": "none (1 GPU)" if world == 1 else
                                   f"per step: ONE NCCL all_gather of logits [B,1000] f32 + kept/assignment indices [B,{dec_cols}] i32 (packed)",
                       "e2e_pipeline": "per step: H2D of the batch (pinned, copy stream, 2 device buffers; overlaps the "
                                       "previous step's forward) + forward + D2H of the logits into one of two pinned buffers; the host waits for "
                                       "step i-1's logits after enqueuing step i (every step's logits are read inside the timed region)"},
            "e2e": {"value": round(total * a.steps / (ms_e2e * 1e-3), 1), "unit": "images/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clock_rec,
            "roofline": None if not top else {
                "kernel": top["kernel"], "bound": "hbm", "achieved": top["alg_gbs"], "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": top["frac_hbm"], "traffic": traffic, "peak_source": peak_kind,
                "avg_us": top["avg_us"], "alg_mb_per_launch": top["alg_mb"]},
            "kernels": kernels,
            "tokred_share_of_step": round(sum(k["avg_us"] * k["launches_per_step"] for k in kernels) / (ms / a.steps * 1e3), 4),
            "workloads": extras,
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
    ap.add_argument("--cpu-batch", type=int, default=0, help="images per step of the CPU arm (0 = the GPU arm's per-GPU batch)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--headline-only", action="store_true", help="skip the other BASELINE configs (workloads array)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying one CUDA graph per step")
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
