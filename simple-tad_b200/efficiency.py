"""Latency / throughput harness — the counterpart of the reference's test_efficiency.py (te:12-196) for the models on
this path, plus the batch-size sweep of BASELINE config 5.

`main(model_type, with_flash)` keeps the reference's call shape and prints the same three lines (average time, FPS,
peak GPU memory) for a batch-1 fp32 clip `x ~ N(0, 1) [1, 3, 16, 224, 224]` (te:17).  Differences, all on the
measurement side: device time is taken with CUDA events around every forward (the reference reads `time.time()` without
a synchronize, te:173-179, i.e. it times the launch), and the forward can be replayed from a CUDA graph
(`use_graph=True`): one batch-1 forward is 87 kernel launches of a few microseconds each, so launch latency is the
bound that a graph removes.

`batch_sweep(...)` runs B = 1 ... 256 clips per GPU (config 5).  Under `torchrun --nproc-per-node N` every rank sweeps
its own GPU (clips are independent: no collective on the data path) and rank 0 reports the per-rank latency together
with the summed clips/s (max over ranks of the device time, one all_reduce per batch size).
"""
import gc
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from . import modeling_finetune  # noqa: F401  (registers the vit_* factories, as `import modeling_finetune` does at te:9)
from .other_models.MVD import modeling_finetune as _mvd  # noqa: F401  (registers mvd_vit_*, te:10)
from .registry import create_model

MODEL_NAMES = {"VideoMAE-S": "vit_small_patch16_224", "VideoMAE-B": "vit_base_patch16_224",
               "VideoMAE-L": "vit_large_patch16_224", "ViViT-B": "vit_base_patch16_224",   # te:22-57, te:96-113 (+ ViT-L)
               "MVD-S": "mvd_vit_small_patch16_224", "MVD-B": "mvd_vit_base_patch16_224"}  # te:58-95


def build(model_type, with_flash=False, device="cuda"):
    """The create_model call of te:24-95, kwarg for kwarg (drop_block_rate=None is filtered by create_model)."""
    if model_type not in MODEL_NAMES:
        raise ValueError(f"{model_type!r}: this path covers {sorted(MODEL_NAMES)} (InternVideo2 is another model "
                         "family: different blocks, patch 14)")
    extra = {"use_cls_token": False} if model_type.startswith("MVD") else {}   # te:72, te:91
    model = create_model(MODEL_NAMES[model_type], pretrained=False, num_classes=2, all_frames=16, tubelet_size=2,
                         fc_drop_rate=0.0, drop_rate=0.0, drop_path_rate=0.1, attn_drop_rate=0.0, drop_block_rate=None,
                         use_checkpoint=False, final_reduction="fc_norm", init_scale=0.001, use_flash_attn=with_flash,
                         **extra)
    return model.to(device).eval()


class GraphedForward:
    """model(x) for a fixed input shape, captured once into a CUDA graph and replayed: `run(x)` copies x into the static
    input and replays.  The library is graph-capturable (no allocation, no synchronisation inside, include/stad.h)."""

    def __init__(self, model, example):
        self.model = model
        self.static_x = example.clone()
        self._capture()

    def _capture(self):
        model = self.model
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up off the capture stream: weight packing, workspace allocation
            for _ in range(2):
                model(self.static_x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = model(self.static_x)
        self._token = model.prepare().graph_token()

    def run(self, x=None):
        # the graph has the prepared model's weight and workspace pointers baked in: re-capture when either has changed
        # since (weights reloaded or moved; the workspace reallocated for a larger batch by another caller)
        if self.model.prepare().graph_token() != self._token:
            self._capture()
        if x is not None and x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


def time_forward(fn, steps, warmup):
    """Per-call device time (ms) of `fn()` over `steps` calls, CUDA events on the current stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def main(model_type, with_flash=False, steps=1000, use_graph=True, quiet=False):
    """te:12-196 for one model type: batch-1 latency, FPS and peak memory; returns them as a dict."""
    gc.collect()
    torch.cuda.empty_cache()
    device = torch.device("cuda")
    model = build(model_type, with_flash, device)
    n_parameters = sum(p.numel() for p in model.parameters() if p.requires_grad)
    x = torch.tensor(np.random.randn(1, 3, 16, 224, 224).astype(np.float32)).to(device)
    with torch.no_grad():
        model(x)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        if use_graph:
            fwd = GraphedForward(model, x)
            times = time_forward(lambda: fwd.run(), steps, warmup=20)
        else:
            times = time_forward(lambda: model(x), steps, warmup=20)
    avg_ms = float(np.mean(times))
    mem = torch.cuda.max_memory_allocated() / (1024 ** 2)
    if not quiet:
        print('Number of params:', n_parameters)
        print(f'{model_type} | Average time: {avg_ms:.2f} ms')
        print(f'{model_type} | Average FPS: {1000.0 / avg_ms:.2f}')
        print(f'{model_type} | Average GPU memory use: {mem:.2f}')
    return {"model": model_type, "params": n_parameters, "avg_ms": avg_ms, "p50_ms": float(np.median(times)),
            "fps": 1000.0 / avg_ms, "peak_mem_mib": mem, "graph": bool(use_graph), "steps": steps}


def batch_sweep(model_type="VideoMAE-B", batches=(1, 2, 4, 8, 16, 32, 64, 128, 256), warmup=20, iters=100,
                use_graph=True, quiet=False):
    """BASELINE config 5: latency and clips/s for B clips per GPU, on every rank of the job."""
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    device = torch.device("cuda", torch.cuda.current_device())
    model = build(model_type, False, device)
    rows = []
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    for B in batches:
        x = torch.randn((B, 3, 16, 224, 224), generator=gen).to(device)
        with torch.no_grad():
            if use_graph:
                fwd = GraphedForward(model, x)
                fn = lambda: fwd.run()  # noqa: E731
            else:
                fn = lambda: model(x)  # noqa: E731
            n_it = iters if B <= 64 else max(10, iters // 4)
            times = time_forward(fn, n_it, warmup if B <= 64 else 5)
        ms = float(np.mean(times))
        worst = torch.tensor([ms], dtype=torch.float64, device=device)
        if distributed:
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        row = {"batch_per_gpu": B, "ms": ms, "ms_max_over_ranks": float(worst), "p50_ms": float(np.median(times)),
               "clips_per_s": world * B / (float(worst) * 1e-3), "n_gpus": world, "graph": bool(use_graph)}
        rows.append(row)
        if rank == 0 and not quiet:
            print(json.dumps(row))
        del x
        if use_graph:
            del fwd
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    if "RANK" in os.environ:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    what = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    if what == "sweep":
        batch_sweep(sys.argv[2] if len(sys.argv) > 2 else "VideoMAE-B")
    else:
        main(what, with_flash=True)
    if dist.is_initialized():
        dist.destroy_process_group()
