"""Multi-GPU check of BASELINE config 3 (run under torchrun on N GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank scores its contiguous block of the windows of `--videos` synthetic videos (ViT-L by default), ONE
all_gather_into_tensor over NCCL returns the full per-frame score table on every rank; rank 0 then scores everything
alone and checks (a) every rank holds the identical table, (b) the sharded table matches the single-GPU one within the
parity tolerance (same kernels; batch composition differs at shard boundaries, so bit-equality holds only per batch)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import synth_data as synth  # noqa: E402  (synthetic weights / videos)
from simple_tad_b200 import modeling_finetune as mf  # noqa: E402
from simple_tad_b200.runner import SlidingWindowRunner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vit_large_patch16_224")
    ap.add_argument("--videos", type=int, default=4)
    ap.add_argument("--frames", type=int, default=40)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = mf.__dict__[a.model](num_classes=2, all_frames=16, tubelet_size=2, init_scale=1.0, final_reduction="fc_norm")
    model.load_state_dict(synth.make_state_dict(a.model, seed=3))
    model = model.to(dev).eval()
    videos = [synth.make_video(a.frames, seed=30 + v) for v in range(a.videos)]
    runner = SlidingWindowRunner(model, batch_windows=16, device=dev)
    table = runner.score_videos(videos)                       # sharded + gathered
    n = a.videos * (a.frames - 15)
    assert table.shape == (n, 2), table.shape
    if world > 1:
        tables = [torch.empty_like(table) for _ in range(world)]
        dist.all_gather(tables, table)
        assert all(torch.equal(t, table) for t in tables), "ranks hold different score tables"
    if rank == 0:
        alone = torch.cat([runner.score_frames_device(v)[0] for v in videos])
        dp = float((alone.softmax(-1) - table.softmax(-1)).abs().max())
        same = float((alone == table).float().mean())
        print(f"dist_check ok: world={world} model={a.model} windows={n}: max|dp| sharded vs single-GPU = {dp:.2e}, "
              f"bit-identical rows = {100 * same:.1f}%")
        assert dp <= 1e-2
    # evaluation epilogue: every rank reduces ITS shard on the device, the ranks all-reduce the 2 x 102 count table
    gen = torch.Generator().manual_seed(11)
    labels = [(torch.rand(v.shape[0], generator=gen) < 0.35).long() for v in videos]
    res, logits2 = runner.evaluate_videos(videos, labels)
    assert torch.equal(logits2, table)
    if rank == 0:
        from simple_tad_b200 import metrics as M
        import numpy as np
        win_labels = torch.cat([l[15:] for l in labels]).to(dev)
        # single-process reference of the same metrics from the gathered table (no process group involved)
        hist, conf = M._lib.eval_hist(table.softmax(-1).contiguous(), win_labels.to(torch.int32), M.threshold_tensor(dev))
        c = M.counts_from_hist(hist.cpu().numpy())
        for k in ("tn", "fp", "fn", "tp"):
            assert np.array_equal(c[k], res["counts"][k]), f"sharded {k} counts differ from the single-process counts"
        assert res["confmat"] == M.argmax_metrics(conf.cpu().numpy())["confmat"] and res["n"] == n
        print(f"dist_check ok: sharded metric counts == single-process counts; auroc {res['auroc']:.4f} ap {res['ap']:.4f} "
              f"acc {res['acc']:.4f} confmat {res['confmat']}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
