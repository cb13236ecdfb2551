"""CPU, world_size 2 over gloo: the clip-sharding and score-gather logic of the multi-GPU runner
(replaces DistributedSampler + gather_predictions_nontensor, rff:311-314 / ut:791-810)."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_covers_everything_once():
    from simple_tad_b200.runner import shard_range
    for n in (0, 1, 7, 85, 680, 681):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_range(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_window_segments_follow_the_dataset_geometry():
    from simple_tad_b200.runner import window_segments
    segs, total = window_segments([100, 20, 15, 16], 0, 10 ** 9)
    assert total == 85 + 5 + 0 + 1                      # T - 15 windows per video (dota.py:209-223)
    assert segs == [(0, 0, 85), (1, 0, 5), (3, 0, 1)]
    segs, _ = window_segments([100, 20, 15, 16], 80, 88)
    assert segs == [(0, 80, 5), (1, 0, 3)]


def _worker(rank, world, port, n_total, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simple_tad_b200.runner import gather_scores, shard_range
    lo, hi, per = shard_range(n_total, world, rank)
    # the score of global window i is (i, -i): any misplacement is visible after the gather
    local = torch.stack([torch.arange(lo, hi, dtype=torch.float32), -torch.arange(lo, hi, dtype=torch.float32)], 1)
    full = gather_scores(local, n_total)
    torch.save(full, os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [85, 680, 7])
def test_gather_scores_world_size_2_gloo(tmp_path, n_total):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    want = torch.stack([torch.arange(n_total, dtype=torch.float32), -torch.arange(n_total, dtype=torch.float32)], 1)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(got, want), f"rank {r} gathered a different score table"


def _metrics_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from oracle import metrics_oracle as mo
    from simple_tad_b200 import metrics as M
    from simple_tad_b200.runner import shard_range
    logits, labels = mo.synthetic_scores(1001, seed=9)
    probs = torch.from_numpy(logits).softmax(-1).numpy()
    lo, hi, _ = shard_range(len(labels), world, rank)
    thr32 = np.asarray(M.THRESHOLDS, dtype=np.float64).astype(np.float32)
    # this rank's shard of the histogram (what stad_eval_hist returns on the device)
    bins = np.searchsorted(thr32, probs[lo:hi, 1], side="right")
    hist = np.zeros((2, 102), dtype=np.int64)
    np.add.at(hist, (labels[lo:hi], bins), 1)
    pred = probs[lo:hi, 1] > probs[lo:hi, 0]
    y = labels[lo:hi].astype(bool)
    conf = np.array([(~pred & ~y).sum(), (pred & ~y).sum(), (~pred & y).sum(), (pred & y).sum()], dtype=np.int64)
    h, c = M.reduce_counts(torch.from_numpy(hist), torch.from_numpy(conf))
    torch.save((h, c), os.path.join(tmp, f"m{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_metric_count_tables_all_reduce_world_size_2_gloo(tmp_path):
    """Each rank reduces its shard, the ranks sum the count tables: every rank ends with the whole-set metrics."""
    import numpy as np
    from oracle import metrics_oracle as mo
    from simple_tad_b200 import metrics as M
    world = 2
    mp.spawn(_metrics_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    logits, labels = mo.synthetic_scores(1001, seed=9)
    probs = torch.from_numpy(logits).softmax(-1).numpy()
    th, counts = mo.thresholded(probs[:, 1], labels, np.asarray(M.THRESHOLDS).astype(np.float32))
    for r in range(world):
        h, c = torch.load(os.path.join(str(tmp_path), f"m{r}.pt"))
        got = M.counts_from_hist(h.numpy())
        assert np.array_equal(got["tp"], counts[:, 3]) and np.array_equal(got["tn"], counts[:, 0])
        assert M.argmax_metrics(c.numpy())["confmat"] == mo.argmax_metrics(probs, labels)["confmat"]


class _IndexModel(torch.nn.Module):
    """Stand-in for the classifier on the host side of the runner: 'scores' a window with (video marker, index of its
    last source frame), read out of the frame tensor itself, so any window mis-addressing shows in the gathered table."""

    def __init__(self, frames_per_clip=16):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.num_frames, self.num_classes = frames_per_clip, 2

    def prepare(self, device=None):
        class _Prepared:  # the real model casts fp32 frames to bf16 once per buffer here; the stand-in keeps them
            def input_bf16(self, x):
                return x
        return _Prepared()

    def forward_windows(self, frames, start=0, count=None, stride=1, frame_step=1, starts=None):
        first = starts.long() if starts is not None else start + torch.arange(count) * stride
        last = first + (self.num_frames - 1) * frame_step
        self.batches = getattr(self, "batches", []) + [int(first.numel())]
        # frame i of video v holds 1000 * v + i in every element: a window that straddled two videos would fail here
        assert torch.equal(frames[first, 0, 0, 0] + (self.num_frames - 1) * frame_step, frames[last, 0, 0, 0])
        lg = torch.stack([frames[last, 0, 0, 0], frames[first, 0, 0, 0]], 1).float()
        return lg, lg.softmax(-1)


def _videos(lengths):
    return [(1000.0 * v + torch.arange(T, dtype=torch.float32)).view(T, 1, 1, 1).expand(T, 3, 2, 2).contiguous()
            for v, T in enumerate(lengths)]


def _runner_worker(rank, world, port, tmp, lengths, stride, frame_step):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simple_tad_b200.runner import SlidingWindowRunner
    runner = SlidingWindowRunner(_IndexModel(), batch_windows=4, device="cpu", stride=stride, frame_step=frame_step)
    full = runner.score_videos(_videos(lengths))
    torch.save(full, os.path.join(tmp, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("stride,frame_step,in_fps", [(1, 1, 10), (5, 3, 30), (10, 1, 10)])
def test_score_videos_windows_follow_the_sequencer_world_size_2_gloo(tmp_path, stride, frame_step, in_fps):
    """The sharded runner scores exactly the windows RegularSequencer lists (dataset/sequencing.py:38-62): end-aligned,
    `stride` source frames apart, `frame_step` source frames between the frames of a window; video by video, in order,
    identical on both ranks."""
    from simple_tad_b200.sequencing import window_plan
    lengths = [58, 50, 46, 40, 100, 16]
    world = 2
    mp.spawn(_runner_worker, args=(world, _free_port(), str(tmp_path), lengths, stride, frame_step), nprocs=world, join=True)
    rows = []
    for v, T in enumerate(lengths):
        plan = window_plan(T, in_fps, 10, 16, stride)
        if plan is None:
            continue
        assert plan.frame_step == frame_step
        rows += [(1000.0 * v + seq[-1], 1000.0 * v + seq[0]) for seq in plan.sequences()]
    want = torch.tensor(rows, dtype=torch.float32)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"s{r}.pt"))
        assert got.shape == want.shape and torch.equal(got, want), f"rank {r}: window table differs from the sequencer's"


@pytest.mark.parametrize("world", [4, 8])
def test_score_videos_config3_geometry_world_size_4_and_8_gloo(tmp_path, world):
    """BASELINE config 3 geometry (8 videos x 100 frames = 680 windows) sharded over 4 and 8 ranks: at 8 ranks every
    shard is 85 windows = one video's worth but NOT aligned to the videos' window ranges once padding enters; every rank
    must end up with the identical, complete [680, 2] table."""
    lengths = [100] * 8
    mp.spawn(_runner_worker, args=(world, _free_port(), str(tmp_path), lengths, 1, 1), nprocs=world, join=True)
    want = torch.tensor([(1000.0 * v + w + 15, 1000.0 * v + w) for v in range(8) for w in range(85)])
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"s{r}.pt"))
        assert got.shape == want.shape and torch.equal(got, want), f"rank {r} of {world}: window table differs"


def test_score_videos_runs_full_batches_across_video_boundaries():
    """BASELINE config 3 shape (8 videos x 100 frames = 8 x 85 windows): the shard of a rank is scored in FULL batches of
    `batch_windows` windows wherever the video boundaries fall (the reference's DataLoader batches windows across
    videos, rff:311-314), in chunks of at most `max_frames` resident frames; the table is the sequencer's."""
    from simple_tad_b200.runner import SlidingWindowRunner
    lengths = [100] * 8
    model = _IndexModel()
    runner = SlidingWindowRunner(model, batch_windows=64, device="cpu")
    full = runner.score_videos(_videos(lengths))
    assert model.batches == [64] * 10 + [40]
    want = torch.tensor([(1000.0 * v + w + 15, 1000.0 * v + w) for v in range(8) for w in range(85)])
    assert torch.equal(full, want)
    # chunking by resident frames: pieces never exceed the budget, a long video is split by windows, nothing is lost
    runner = SlidingWindowRunner(_IndexModel(), batch_windows=16, device="cpu", stride=2)
    lengths = [300, 16, 40, 17]
    segs = [(0, 0, 143), (1, 0, 1), (2, 0, 13), (3, 0, 1)]
    chunks = runner.plan_chunks(lengths, segs, max_frames=64)
    n_windows = 0
    for chunk in chunks:
        used = 0
        for v, f0, f1, off, ws in chunk:
            assert off == used and 0 <= f0 < f1 <= lengths[v]
            assert all(off <= w and w + 16 <= off + (f1 - f0) for w in ws)
            used += f1 - f0
            n_windows += len(ws)
        assert used <= 64
    assert n_windows == sum(c for _, _, c in segs)
    got = runner._score_segments(_videos(lengths), lengths, segs, max_frames=64)
    want = []
    for v, T in enumerate(lengths):
        first = (T - 16) % 2
        want += [(1000.0 * v + first + 2 * w + 15, 1000.0 * v + first + 2 * w) for w in range((T - 16) // 2 + 1)]
    assert torch.equal(got, torch.tensor(want))
