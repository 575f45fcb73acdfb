"""CPU checks of host-side logic and of exactness arguments the kernels rely on (no GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _argmin_sqrt_brute(x):
    d = np.sqrt(np.maximum(x, np.float32(1e-30))).astype(np.float32)
    best, bp = np.float32(np.inf), 0
    for p, v in enumerate(d):
        if v < best:
            best, bp = v, p
    return bp


def _argmin_sqrt_walk(x):
    """csrc/ats.cu's sampling step: minimum radicand, ulp walk to the largest float with the same square root, first index."""
    xc = np.maximum(x, np.float32(1e-30)).astype(np.float32)
    xmin = xc.min()
    if not xmin < np.inf:
        return 0
    dmin = np.sqrt(xmin).astype(np.float32)
    big = xmin
    for _ in range(8):
        nxt = np.frombuffer(np.uint32(np.frombuffer(np.float32(big).tobytes(), np.uint32)[0] + 1).tobytes(), np.float32)[0]
        if not (np.sqrt(nxt).astype(np.float32) == dmin):
            break
        big = nxt
    idx = np.nonzero(xc <= big)[0]
    return int(idx[0]) if len(idx) else 0


def test_ats_argmin_without_sqrt_per_entry_is_exact():
    """sqrtf is monotone and at most three floats share a square root: the first entry whose radicand is <= the largest
    float with sqrt == sqrt(min) is the lowest-index argmin of sqrt(max(x, 1e-30)) -- checked on adversarial inputs whose
    radicands sit within a few ulps of each other (correctly rounded sqrt on both sides, like CUDA's sqrtf)."""
    rng = np.random.default_rng(0)
    for trial in range(4000):
        n = int(rng.integers(2, 40))
        base = np.float32(rng.uniform(1e-6, 4.0)) if trial % 3 else np.float32(rng.uniform(1e-12, 1e-3))
        bits = np.frombuffer(np.float32(base).tobytes(), np.uint32)[0]
        x = np.frombuffer((bits + rng.integers(0, 6, size=n).astype(np.uint32)).astype(np.uint32).tobytes(), np.float32).copy()
        if trial % 5 == 0:
            x[rng.integers(0, n)] *= np.float32(1.5)
        if trial % 7 == 0:
            x[rng.integers(0, n)] = np.float32(0.0)          # clamps to 1e-30
        assert _argmin_sqrt_brute(x) == _argmin_sqrt_walk(x), (trial, x)


def test_bench_numa_binding_is_a_noop_without_gpus():
    """bench.bind_to_gpu_numa_node never raises: without CUDA / NVML it leaves the affinity alone and says so."""
    import torch
    import bench
    before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    assert bench.bind_to_gpu_numa_node(torch.device("cuda", 0)) is None
    if before is not None:
        assert os.sched_getaffinity(0) == before
    os.environ["TOKRED_BENCH_NO_AFFINITY"] = "1"
    try:
        assert bench.bind_to_gpu_numa_node(torch.device("cuda", 0)) is None
    finally:
        del os.environ["TOKRED_BENCH_NO_AFFINITY"]
