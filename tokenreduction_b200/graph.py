"""CUDA-graph replay of an inference forward.

A reduced-ViT forward is 150-450 launches of 5-100 us; with the fused attention producer the device time of a DeiT-S
step dropped below the time Python needs to ISSUE those launches, so the step became CPU-bound.  Every tokred entry point
only enqueues on the current stream (no allocation, no synchronisation, no host read -- include/tokred.h), and the models
make no host-side decision on device data (ATS pads to its static width), so a whole forward captures into one CUDA graph
and a step becomes ONE launch.

    run = GraphedForward(lambda x: model(x), example_batch, autocast_dtype=torch.bfloat16)
    logits = run(batch)            # copies `batch` into the static input (skipped when batch IS run.static_input)

Outputs are the graph's static tensors: valid until the next call.
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch


class GraphedForward:
    def __init__(self, fn: Callable[[torch.Tensor], Any], example: torch.Tensor, autocast_dtype: Optional[torch.dtype] = None,
                 static_input: Optional[torch.Tensor] = None, warmup: int = 2):
        if not example.is_cuda:
            raise RuntimeError("GraphedForward needs a CUDA example input")
        self.fn, self.autocast_dtype = fn, autocast_dtype
        self.static_input = static_input if static_input is not None else example.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the capture: lazy module loads, cuBLAS workspaces, autotuning
            for _ in range(warmup):
                self._call(self.static_input)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_output = self._call(self.static_input)

    def _call(self, x):
        with torch.no_grad(), torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16,
                                            enabled=self.autocast_dtype is not None):
            return self.fn(x)

    def __call__(self, x: torch.Tensor):
        if x.data_ptr() != self.static_input.data_ptr():
            self.static_input.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_output
