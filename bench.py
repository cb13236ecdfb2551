"""Benchmark of the hot path: ViT-B/16 16x224^2 frame-level sliding-window inference (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One "step" = one batch of `--batch` (64) stride-1 windows of a synthetic DoTA-shaped video scored through
VisionTransformer.forward_windows (fp32->bf16 cast, patch embed, 12 blocks, pool/norm/head): clips/s = windows/s.
Prints ONE JSON line (rank 0).  Keys follow the driver contract; see DESIGN.md "Measurement".
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

GFLOP_PER_CLIP = {  # F = 2*N*1536*D + L*(24*N*D^2 + 4*N^2*D), N = 1568
    "vit_small_patch16_224": 113.76, "vit_base_patch16_224": 360.69, "vit_large_patch16_224": 1193.67,
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="stad", choices=["stad", "reference"])
    ap.add_argument("--model", default="vit_base_patch16_224", choices=sorted(GFLOP_PER_CLIP))
    ap.add_argument("--batch", type=int, default=64, help="windows (clips) per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-clips", type=int, default=1, help="--impl reference: clips per step (bounded sample)")
    ap.add_argument("--preheat-s", type=float, default=2.5, help="untimed steps run for this long before the timed region "
                                                                  "(the SM clock settles under the power cap)")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-3 / config-5 legs (keys c3, c5)")
    return ap.parse_args()


def load_ncu_traffic(model_name, B):
    """DRAM bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    capture of this workload (profiles/ncu_gemm_dram.json, written by tools/ncu_dram_summary.py); None when the capture
    is for another model / batch."""
    path = os.path.join(ROOT, "profiles", "ncu_gemm_dram.json")
    if not os.path.exists(path):
        return None
    d = json.load(open(path))
    if d.get("model") != model_name or d.get("batch") != B:
        return None
    return d


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(model_name, B, world, sample=None):
    """`config` of both arms: the workload BASELINE.json's metric is quoted on (config 2)."""
    import synth_data as synth
    D = synth.ARCHS[model_name][0]
    cfg = {"workload": f"{model_name} 16x224^2 frame-level sliding-window inference, {B} stride-1 windows "
                       f"(one {B + 15}-frame DoTA-shaped synthetic video chunk) per step per GPU, 2-class head",
           "batch_per_gpu": B, "parallelism": f"clip-sharded x{world}, one score gather",
           "l2": f"inputs rotate over 4 distinct videos; activations per step ({B * 1568 * D * 2 * 5 / 1e6:.0f} MB) "
                 "exceed the 126 MB L2"}
    if sample:
        cfg["sample"] = sample
    return cfg


def cpu_forward_setup(model_name):
    """The reference's CPU forward: the UNMODIFIED modeling_finetune.VisionTransformer from oracle/_ref/ when the build
    step put it there (kind "reference"), else oracle/vit_oracle.py, its restatement pinned by tests/golden (kind "port").
    Returns (run(clips) -> logits, synth module, kind)."""
    from oracle import ref_loader, synth, vit_oracle
    D, depth, heads = synth.ARCHS[model_name]
    sd = synth.make_state_dict(model_name, seed=0)
    torch.set_num_threads(os.cpu_count())
    mf_ref = ref_loader.load()
    if mf_ref is not None and model_name in mf_ref.__dict__:
        model = mf_ref.__dict__[model_name](num_classes=2, all_frames=16, tubelet_size=2, use_flash_attn=False,
                                            init_scale=1.0, final_reduction="fc_norm")
        model.load_state_dict(sd, strict=True)
        model.eval()

        @torch.no_grad()
        def run_ref(clips):
            return model(clips)
        return run_ref, synth, "reference"

    def run(clips):
        return vit_oracle.vit_forward(sd, clips, heads)
    return run, synth, "port"


def cpu_baseline(model_name, clips=8, repeats=5):
    """Bounded sample of the workload on the host cores: ~10 s of CPU work (5 forwards of 8 clips at 3-6 clips/s)."""
    run, synth, kind = cpu_forward_setup(model_name)
    x = synth.make_clips(clips, seed=123)
    run(x[:1])  # warm-up (thread pool, oneDNN primitives)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        run(x)
        best = min(best, time.perf_counter() - t0)
    what = ("the unmodified reference modeling_finetune.py (oracle/_ref)" if kind == "reference" else
            "oracle/vit_oracle.py (the port)")
    return {"value": clips / best, "unit": "clips/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"best of {repeats} fp32 forwards of {clips} synthetic clips ({model_name}) through {what} "
                      f"(torch {torch.__version__} CPU, {torch.get_num_threads()} threads)"}


def run_reference(args, rank, world, out):
    """--impl reference: the reference's own CPU implementation of the path — the unmodified modeling_finetune.py copied
    into oracle/_ref/ by the build step (it travels to the GPU box with the snapshot), else oracle/ (its restatement,
    checked against the reference's outputs in tests/golden).  Rank 0 alone runs it."""
    if rank != 0:
        return
    run, synth, kind = cpu_forward_setup(args.model)
    x = synth.make_clips(args.ref_clips, seed=123)
    for _ in range(max(1, min(args.warmup, 2))):
        run(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(x)
    dt = time.perf_counter() - t0
    value = args.ref_clips * args.steps / dt
    line = {
        "impl": "reference", "metric": "clips/sec (16x224^2 bf16)", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.model, args.batch, args.gpus,
                                  sample=f"bounded sample: {args.ref_clips} clip(s) of the workload per step, fp32, on the "
                                         f"host CPU ({torch.get_num_threads()} threads); rank 0 only"),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": os.cpu_count(), "kind": kind,
                         "sample": f"{args.steps} steps x {args.ref_clips} clip(s), "
                                   f"{'unmodified reference modeling_finetune.py (oracle/_ref)' if kind == 'reference' else 'oracle/vit_oracle.py'}"
                                   f" fp32, {torch.get_num_threads()} threads"},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.emit(json.dumps(line))


def config3_leg(dev, rank, world, synth, mf):
    """BASELINE config 3 as an extra key of the bench line (the headline stays config 2).
    c3: ViT-L/16, 8 synthetic 100-frame videos = 8 x 85 windows through SlidingWindowRunner.score_videos, the window
        index space sharded over the ranks (STRONG scaling) with one score gather; the gathered [680, 2] table is compared
        bit for bit with the same table computed by one rank alone inside this job (rff:311-314, eff:449-454, ut:791-810)."""
    from simple_tad_b200.runner import SlidingWindowRunner
    name = "vit_large_patch16_224"
    model = mf.__dict__[name](num_classes=2, all_frames=16, tubelet_size=2, init_scale=1.0, final_reduction="fc_norm",
                              use_flash_attn=True)
    model.load_state_dict(synth.make_state_dict(name, seed=3))
    model = model.to(dev).eval()
    videos = [synth.make_video(100, seed=300 + v).pin_memory() for v in range(8)]
    runner = SlidingWindowRunner(model, batch_windows=64, device=dev)
    runner.score_videos(videos[:1])  # warm-up: weight packing, workspace
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        table = runner.score_videos(videos)  # sharded + gathered; frames go host -> device inside
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c3 = {"workload": "vit_large_patch16_224, 8 videos x 100 frames = 680 windows, clip-sharded, one score gather",
          "scaling": "strong", "n_gpus": world, "windows": int(table.shape[0]), "clips_per_s": table.shape[0] / float(t),
          "seconds": float(t), "host_frames_in": True}
    if world > 1:
        # the same table by ONE rank alone (no process group involved): every rank checks its gathered copy against it
        lengths = [int(v.shape[0]) for v in videos]
        alone = None
        if rank == 0:
            t0 = time.perf_counter()
            alone = runner._score_segments(videos, lengths, [(v, 0, lengths[v] - 15) for v in range(8)])
            torch.cuda.synchronize()
            c3["single_rank_seconds"] = time.perf_counter() - t0
        else:
            alone = torch.empty_like(table)
        dist.broadcast(alone, src=0)
        same = torch.tensor([int(torch.equal(alone, table))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        c3["gathered_table_bit_identical_to_single_rank"] = bool(int(same))
        # batches of another size pick other GEMM tile widths, hence another grouping of the LayerNorm partial sums: a
        # shard whose ragged last batch differs from the single rank's can differ in the last bits; report by how much
        diff = (alone.double() - table.double()).abs().max().reshape(1)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        c3["max_abs_diff_vs_single_rank"] = float(diff)
        if rank == 0:
            c3["efficiency_vs_single_rank"] = c3["single_rank_seconds"] / (world * c3["seconds"])
    del model, runner, videos
    torch.cuda.empty_cache()
    return c3


def config5_leg(world):
    """c5: ViT-B graph-replay batch sweep B = 1 ... 256 clips per GPU on every rank (test_efficiency.py shape, te:174-194)."""
    from simple_tad_b200 import efficiency
    rows = efficiency.batch_sweep("VideoMAE-B", batches=(1, 2, 4, 8, 16, 32, 64, 128, 256), warmup=10, iters=40, quiet=True)
    return {"workload": "vit_base_patch16_224 CUDA-graph replay, B clips per GPU resident in HBM (fp32 in), per rank",
            "n_gpus": world, "rows": [{k: r[k] for k in ("batch_per_gpu", "ms_max_over_ranks", "clips_per_s")} for r in rows]}


class OnlyJsonOnStdout:
    """Everything written to fd 1 while the bench runs (NCCL's version banner, library chatter) goes to stderr; the
    driver reads ONE JSON line from stdout, printed through emit()."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


def main():
    args = parse_args()
    out = OnlyJsonOnStdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    import synth_data as synth  # synthetic weights / inputs (repo-level module, not part of oracle/)
    from simple_tad_b200 import _lib, modeling_finetune as mf
    from simple_tad_b200.runner import SlidingWindowRunner, gather_scores

    D, depth, heads = synth.ARCHS[args.model]
    model = mf.__dict__[args.model](num_classes=2, all_frames=16, tubelet_size=2, init_scale=1.0,
                                    final_reduction="fc_norm", use_flash_attn=True)
    model.load_state_dict(synth.make_state_dict(args.model, seed=0))
    model = model.to(dev).eval()
    prep = model.prepare(dev)

    B = args.batch
    T_frames = B + 15  # B stride-1 windows need B + 15 frames
    n_bufs = 4         # rotate distinct videos so the inputs of a step are never L2-resident from the step before
    host_videos = [synth.make_video(T_frames, seed=10 * rank + i).pin_memory() for i in range(n_bufs)]
    dev_videos = [v.to(dev) for v in host_videos]
    runner = SlidingWindowRunner(model, batch_windows=B, device=dev)
    scores = torch.zeros(args.steps * B, 2, dtype=torch.float32, device=dev)

    def step(i, out=None):
        lg, pr = model.forward_windows(dev_videos[i % n_bufs], start=0, count=B)
        if out is not None:
            out[i * B:(i + 1) * B] = pr
        return lg

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident throughput (`value`)
    for i in range(args.warmup):
        step(i)
    launches_per_step = prep.last_launches + 1  # + the fp32->bf16 cast of the frames
    # untimed pre-heat: the part settles to its power-capped clock within ~2 s of dense tensor work; the timed region
    # then sees the sustained clock the roofline denominator (bf16_tflops_sustained) was measured at
    torch.cuda.synchronize()
    t_heat, n_heat = time.perf_counter(), 0
    while time.perf_counter() - t_heat < args.preheat_s:
        for i in range(4):
            step(i)
        torch.cuda.synchronize()
        n_heat += 4
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i, scores)
    if world > 1:
        gather_scores(scores, world * scores.shape[0])  # the per-frame score gather (NCCL over NVLink), once
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

    # ---------------------------------------------------------------- same steps with per-launch events (roofline)
    cap = (launches_per_step + 4) * args.steps
    _lib.profile_enable(cap)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(args.steps):
        step(i)
    p1.record()
    torch.cuda.synchronize()
    records = _lib.profile_read(cap)
    _lib.profile_enable(0)
    ms_profiled = p0.elapsed_time(p1) / args.steps
    clocks = sampler.stop()

    by_kind = {}
    for kind, epi, m, n, k, ms in records:
        if kind == "gemm":
            flops = 2.0 * m * n * k
        elif kind == "attention":
            flops = 4.0 * m * n * float(k) * k * 64
        else:
            flops = 0.0
        d = by_kind.setdefault(kind, {"ms": 0.0, "flops": 0.0, "launches": 0})
        d["ms"] += ms
        d["flops"] += flops
        d["launches"] += 1
    peaks = load_peaks()
    total_kernel_ms = sum(d["ms"] for d in by_kind.values()) or 1.0
    gemm = by_kind.get("gemm", {"ms": 1.0, "flops": 0.0, "launches": 1})
    att = by_kind.get("attention", {"ms": 1.0, "flops": 0.0, "launches": 1})
    gemm_tflops = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
    att_tflops = att["flops"] / (att["ms"] * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]  # kernels timed inside a long step -> sustained figure
    roofline = {
        "bound": "tensor", "kernel": "gemm_kernel (tcgen05, all epilogues)", "achieved": gemm_tflops, "peak": peak,
        "unit": "TFLOP/s", "frac": gemm_tflops / peak, "traffic": None, "traffic_unit": "bytes per launch",
        "peak_source": peaks["source"] + ", sustained cuBLAS bf16",
        "launches_per_step": gemm["launches"] / args.steps, "avg_launch_ms": gemm["ms"] / max(1, gemm["launches"]),
        "share_of_step": gemm["ms"] / total_kernel_ms,
        "kernel_ms_per_step": total_kernel_ms / args.steps,  # sum of all event-timed launches; the rest is launch gaps
        "attention": {"achieved": att_tflops, "frac": att_tflops / peak, "share_of_step": att["ms"] / total_kernel_ms,
                      "avg_launch_ms": att["ms"] / max(1, att["launches"])},
        "other_share_of_step": {k: v["ms"] / total_kernel_ms for k, v in by_kind.items() if k not in ("gemm", "attention")},
    }

    ncu = load_ncu_traffic(args.model, B)
    if ncu is not None:
        # weighted by how often each GEMM shape launches in a step (1 patch embed + depth x {qkv, proj, fc1, fc2})
        per = ncu["dram_bytes_per_launch"]
        tot = per["patch_embed"] + depth * (per["qkv"] + per["proj"] + per["fc1"] + per["fc2"])
        roofline["traffic"] = tot / (1 + 4 * depth)
        alg = ncu["algorithmic_bytes_per_launch"]
        roofline["traffic_detail"] = {"dram_bytes_per_launch": per, "algorithmic_bytes_per_launch": alg,
                                      "source": ncu["source"]}
        roofline["traffic_source"] = ("NOT measured in this run (ncu cannot run inside the timed bench): read from "
                                      "profiles/ncu_gemm_dram.json, captured " + str(ncu.get("captured", "in round 1")))

    # ---------------------------------------------------------------- end to end through the public runner API
    # Host frames as the reference's callers hold them after cv2.resize (uint8 HWC BGR, ri:79-81): every step copies
    # them host->device from pinned memory, normalises on the device (prepare_image, ri:15-34), scores the B windows
    # and reads (logits, probs) back to the host.
    gen = torch.Generator().manual_seed(77 + rank)
    host_u8 = [torch.randint(0, 256, (T_frames, 224, 224, 3), generator=gen, dtype=torch.uint8).pin_memory()
               for _ in range(n_bufs)]
    e2e_steps = max(3, args.steps)
    runner.score_frames_u8(host_u8[0], bgr=True)
    sync_all()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        lg, pr = runner.score_frames_u8(host_u8[i % n_bufs], bgr=True)  # pinned host frames in, host scores out
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = host_u8[0].numel() * host_u8[0].element_size()
    d2h = 2 * B * 2 * 4
    e2e_launches = prep.last_launches + 1  # + the uint8 normalise kernel

    clips_per_s = world * B * args.steps / (ms_total * 1e-3)
    model_tflops = clips_per_s / world * GFLOP_PER_CLIP[args.model] / 1e3
    line = {
        "metric": "clips/sec (16x224^2 bf16)", "value": clips_per_s, "unit": "clips/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args.model, B, world),
        "model_tflops_per_gpu": model_tflops,
        "model_frac_of_peak": {"measured_sustained": model_tflops / peaks["bf16_tflops_sustained"],
                               "measured_burst": model_tflops / peaks["bf16_tflops"], "spec_2250": model_tflops / 2250.0},
        "ms_per_step_profiled": ms_profiled,
        "roofline": roofline,
        "clocks": clocks,
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "clips/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "gpu_launches_per_step": e2e_launches,
                "api": "SlidingWindowRunner.score_frames_u8(pinned uint8 HWC BGR frames) -> host (logits, probs)"},
        "gpu_launches": launches_per_step * args.steps,
        "preheat": {"seconds": args.preheat_s, "untimed_steps": n_heat},
        "notes": ["`value` (frames resident in HBM as fp32) includes the fp32->bf16 cast of the step's frames "
                  f"({T_frames * 3 * 224 * 224 * 4 / 1e6:.0f} MB read); `e2e` starts from pinned uint8 host frames (H2D copy + "
                  "uint8 normalise kernel instead of the cast) and ends with the scores on the host: the two legs do not time "
                  "the same work, so e2e can exceed value",
                  "the reference arm (--impl reference) runs a bounded sample (--ref-clips clips per step) of this workload"],
    }
    if not args.no_extras:
        # the extra legs never take the headline down with them: a failure (the same on every rank, e.g. out of memory)
        # is reported under the leg's own key and the line above stands
        for key, leg in (("c3", lambda: config3_leg(dev, rank, world, synth, mf)), ("c5", lambda: config5_leg(world))):
            try:
                line[key] = leg()
            except Exception as e:  # noqa: BLE001
                line[key] = {"error": f"{type(e).__name__}: {e}"[:400]}
                torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.model)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        out.emit(json.dumps(line))


if __name__ == "__main__":
    main()
