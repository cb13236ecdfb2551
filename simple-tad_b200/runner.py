"""Frame-level sliding-window inference (the callers of the hot path): run_inference.py:69-109,
run_inference_simple.py:428-465 (streaming one video) and engine_for_frame_finetuning.final_test:385-463 with
utils.gather_predictions_nontensor:791-810 (a dataset of videos sharded across the GPUs of one node).

B200-first differences from the reference, none of which change the scores:
  * windows are never materialised: frames are uploaded once per video chunk and the patch-embed kernel reads every
    window straight out of the frame buffer (15/16 of the bytes of consecutive windows are shared);
  * ranks take CONTIGUOUS blocks of the window index space (DistributedSampler interleaves, rff:311-314), so a rank
    uploads each frame at most once;
  * one fixed-size all_gather_into_tensor of fp32 scores replaces the pickle-based all_gather_object (ut:803).
Clips are independent, so there is no collective on the data path."""
import math

import torch
import torch.distributed as dist

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # model.default_cfg mean / std, ri:56-62
IMAGENET_STD = (0.229, 0.224, 0.225)


def shard_range(n, world_size, rank):
    """Contiguous block of window indices owned by `rank`: (lo, hi, per_rank) with per_rank = ceil(n / world)."""
    per = max(1, math.ceil(n / max(1, world_size)))
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi, per


def gather_scores(local, n_total, group=None):
    """Every rank contributes its [hi-lo, C] block (padded to per_rank rows); returns the full [n_total, C] tensor on
    every rank.  Works on CUDA tensors over NCCL (NVLink) and on CPU tensors over gloo (tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local[:n_total]
    world = dist.get_world_size(group)
    per = max(1, math.ceil(n_total / world))
    C = local.shape[1]
    padded = torch.zeros(per, C, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty(world * per, C, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_total]


def window_segments(video_lengths, lo, hi, frames_per_clip=16, stride=1, frame_step=1):
    """Map the global window range [lo, hi) onto (video, first_window, n_windows) segments.
    A video of T frames has (T - span) // stride + 1 windows, span = (frames_per_clip - 1) * frame_step + 1
    (sequencing.py:38-62, simple_tad_b200.sequencing.window_plan)."""
    segs = []
    base = 0
    span = (frames_per_clip - 1) * frame_step + 1
    for v, T in enumerate(video_lengths):
        n = (T - span) // stride + 1 if T >= span else 0
        a, b = max(lo, base), min(hi, base + n)
        if a < b:
            segs.append((v, a - base, b - a))
        base += n
    return segs, base


class SlidingWindowRunner:
    """Scores every 16-frame window of one or many videos with a VisionTransformer of this package."""

    def __init__(self, model, batch_windows=64, device=None, stride=1, frame_step=1):
        """stride: source frames between consecutive windows (`view_step`, dota.py:209); frame_step: source frames
        between consecutive frames of a window (orig_fps // target_fps: 3 for DADA-2000, dada.py:31).  Windows are
        aligned to the end of a video, as RegularSequencer aligns them (sequencing.py:53-60)."""
        self.model = model.eval()
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.batch_windows = int(batch_windows)
        self.stride = int(stride)
        self.frame_step = int(frame_step)
        self.T = model.num_frames
        self.span = (self.T - 1) * self.frame_step + 1
        self._dev_frames = None
        self._norm_frames = None

    def _upload(self, frames):
        """Host frames [F, C, H, W] -> reused device buffer (async copy on the current stream if pinned)."""
        if frames.is_cuda:
            return frames
        n = frames.numel()
        if self._dev_frames is None or self._dev_frames.numel() < n or self._dev_frames.dtype != frames.dtype:
            self._dev_frames = torch.empty(n, dtype=frames.dtype, device=self.device)
        dst = self._dev_frames[:n].view(frames.shape)
        dst.copy_(frames, non_blocking=True)
        return dst

    @torch.no_grad()
    def score_frames_device(self, frames):
        """frames [F, C, H, W] (host or device) -> (logits, probs) [(F - span) // stride + 1, num_classes] on the device."""
        dev = self._upload(frames)
        F_ = dev.shape[0]
        n = (F_ - self.span) // self.stride + 1
        if F_ < self.span:
            raise ValueError(f"need at least {self.span} frames, got {F_}")
        first = (F_ - self.span) % self.stride      # the last window ends on the last frame (sequencing.py:55)
        # fp32 frames are cast to bf16 ONCE per buffer here, not once per batch of windows inside forward_windows
        dev = self.model.prepare(dev.device).input_bf16(dev)
        logits, probs = [], []
        for w0 in range(0, n, self.batch_windows):
            cnt = min(self.batch_windows, n - w0)
            lg, pr = self.model.forward_windows(dev, start=first + w0 * self.stride, count=cnt, stride=self.stride,
                                                frame_step=self.frame_step)
            if n > self.batch_windows:  # the model returns views of reused output buffers
                lg, pr = lg.clone(), pr.clone()
            logits.append(lg)
            probs.append(pr)
        return (logits[0], probs[0]) if len(logits) == 1 else (torch.cat(logits), torch.cat(probs))

    @torch.no_grad()
    def score_frames(self, frames):
        """As above, returning host tensors: the per-frame anomaly probabilities a caller prints (ri:104-108).
        Window w's score belongs to frame w + T - 1 (label = last frame, dota.py:217-223)."""
        logits, probs = self.score_frames_device(frames)
        return logits.cpu(), probs.cpu()

    @torch.no_grad()
    def score_frames_u8(self, frames_u8, bgr=True, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        """uint8 frames [F, H, W, 3] exactly as cv2 hands them over after the resize (BGR when bgr=True, ri:79-81)
        -> host (logits, probs).  The H2D copy moves 1 byte per value (the fp32 path of the reference moves 4, ri:82)
        and prepare_image's cvtColor / div 255 / normalise / HWC->CHW (ri:15-34) run in one kernel on the device
        (stad_normalize_frames_u8) straight into the bf16 frame buffer the patch-embed kernel reads."""
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
            raise ValueError(f"expected uint8 frames [F, H, W, 3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        dev = self._upload(frames_u8.contiguous())
        n = dev.numel()
        if self._norm_frames is None or self._norm_frames.numel() < n:
            self._norm_frames = torch.empty(n, dtype=torch.bfloat16, device=self.device)
        F_, H, W, _ = dev.shape
        size = tuple(self.model.patch_embed.img_size)
        if (H, W) != size:
            # frames as the camera / decoder delivers them (e.g. 720 x 1280): cv2.resize(..., INTER_CUBIC) of ri:79-80 on
            # the device, in OpenCV's fixed-point arithmetic
            from . import frames as _frames
            dev = _frames.resize_cubic_u8(dev, size)
            H, W = size
            n = dev.numel()
            if self._norm_frames is None or self._norm_frames.numel() < n:
                self._norm_frames = torch.empty(n, dtype=torch.bfloat16, device=self.device)
        planes = _lib.normalize_frames_u8(dev, mean, std, bgr=bgr, out=self._norm_frames[:n].view(F_, 3, H, W))
        logits, probs = self.score_frames_device(planes)
        return logits.cpu(), probs.cpu()

    def plan_chunks(self, lengths, segs, max_frames=2048):
        """Lay the frame ranges of the segments (video, first_window, n_windows) end to end in chunks of at most
        `max_frames` frames (a segment longer than that is split by windows).  Returns a list of chunks; a chunk is a list
        of (video, f0, f1, offset, window_starts) with [f0, f1) the source frames copied to chunk frames [offset,
        offset + f1 - f0) and window_starts the chunk-relative first frame of every window of the piece."""
        chunks, cur, used = [], [], 0
        per_piece = max(1, (max_frames - self.span) // self.stride + 1)  # windows whose frames fit one chunk
        for v, w0, cnt in segs:
            first = (lengths[v] - self.span) % self.stride  # windows are aligned to the end of a video (sequencing.py:55)
            done = 0
            while done < cnt:
                take = min(cnt - done, per_piece)
                f0 = first + (w0 + done) * self.stride
                f1 = first + (w0 + done + take - 1) * self.stride + self.span
                if cur and used + (f1 - f0) > max_frames:
                    chunks.append(cur)
                    cur, used = [], 0
                cur.append((v, f0, f1, used, [used + i * self.stride for i in range(take)]))
                used += f1 - f0
                done += take
        if cur:
            chunks.append(cur)
        return chunks

    @torch.no_grad()
    def _score_segments(self, videos, lengths, segs, max_frames=2048):
        """Logits [sum of n_windows, C] (device) of the segments, in order.  The frames of a chunk of segments are
        uploaded once into one device buffer and every batch holds `batch_windows` windows regardless of where the
        video boundaries fall (explicit window starts, stad_input.window_starts): a rank's shard runs FULL batches, as
        the reference's DataLoader batches windows across videos (rff:311-314, eff:418-431)."""
        C = self.model.num_classes
        outs = []
        prep = None
        for chunk in self.plan_chunks(lengths, segs, max_frames):
            n_frames = chunk[-1][3] + chunk[-1][2] - chunk[-1][1]
            sample = videos[chunk[0][0]]
            shape = (n_frames,) + tuple(sample.shape[1:])
            numel = n_frames * sample[0].numel()
            if self._dev_frames is None or self._dev_frames.numel() < numel or self._dev_frames.dtype != sample.dtype:
                self._dev_frames = torch.empty(numel, dtype=sample.dtype, device=self.device)
            buf = self._dev_frames[:numel].view(shape)
            starts = []
            for v, f0, f1, off, ws in chunk:
                buf[off:off + f1 - f0].copy_(videos[v][f0:f1], non_blocking=True)
                starts.extend(ws)
            starts = torch.tensor(starts, dtype=torch.int32).to(self.device, non_blocking=True)
            prep = prep or self.model.prepare(self.device)
            planes = prep.input_bf16(buf)
            for i in range(0, starts.numel(), self.batch_windows):
                st = starts[i:i + self.batch_windows]
                lg, _ = self.model.forward_windows(planes, starts=st, frame_step=self.frame_step)
                outs.append(lg.clone())
        return torch.cat(outs) if outs else torch.zeros(0, C, dtype=torch.float32, device=self.device)

    @torch.no_grad()
    def score_videos(self, videos, group=None):
        """videos: list of host frame tensors [T_v, C, H, W].  The global window index space is split into contiguous
        per-rank blocks; each rank scores its block and ONE gather returns logits [n_windows, num_classes] (same on
        every rank), ordered video by video, window by window."""
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        lengths = [int(v.shape[0]) for v in videos]
        _, n_total = window_segments(lengths, 0, 0, self.T, self.stride, self.frame_step)
        lo, hi, per = shard_range(n_total, world, rank)
        segs, _ = window_segments(lengths, lo, hi, self.T, self.stride, self.frame_step)
        local = self._score_segments(videos, lengths, segs)
        return gather_scores(local, n_total, group)

    @torch.no_grad()
    def evaluate_videos(self, videos, frame_labels, group=None):
        """`final_test` (eff:385-497) for a list of videos: frame_labels[v] is the per-frame 0/1 label tensor [T_v] of
        video v; window w of a video is labelled by its last frame (dota.py:217-223).  Returns (metrics, logits):
        the metrics of simple_tad_b200.metrics.evaluate over ALL windows (each rank reduces its own shard on the device,
        the ranks all-reduce the 2 x 102 count table) and the gathered logits [n_windows, num_classes] for
        predictions.csv."""
        from . import metrics as M
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        lengths = [int(v.shape[0]) for v in videos]
        _, n_total = window_segments(lengths, 0, 0, self.T, self.stride, self.frame_step)
        lo, hi, per = shard_range(n_total, world, rank)
        segs, _ = window_segments(lengths, lo, hi, self.T, self.stride, self.frame_step)
        local = self._score_segments(videos, lengths, segs)
        labs = []
        for v, w0, cnt in segs:
            first = (lengths[v] - self.span) % self.stride
            last = first + torch.arange(w0, w0 + cnt) * self.stride + self.span - 1
            labs.append(torch.as_tensor(frame_labels[v])[last].to(torch.int32))
        local_labels = (torch.cat(labs) if labs else torch.zeros(0, dtype=torch.int32)).to(self.device)
        res = M.evaluate(local.softmax(-1), local_labels, group=group)
        return res, gather_scores(local, n_total, group)


class StreamingScorer:
    """Frame-by-frame scoring of a live video: the loop of run_inference.py:69-109 (fill a 16-frame window, then for
    every new frame drop the oldest, append the newest, predict) without shifting or re-uploading the window.

    The normalised bf16 frames live in a device buffer of 2 x T slots; frame n is written to slots n % T and n % T + T,
    so the last T frames are always one contiguous run [p + 1, p + T] (p = n % T) that the patch-embed kernel reads
    straight out of the buffer (stad_input STAD_IN_FRAMES).  `push` takes the frame exactly as `cv2.resize` returns it
    (uint8 [H, W, 3], BGR), uploads 150 KB, normalises it on the device (prepare_image, ri:15-34) and - once T frames
    have arrived - returns (logits [num_classes], probs [num_classes]) of the window ending at this frame.
    `use_graphs=True` replays one captured CUDA graph per buffer phase (T graphs, captured lazily): a batch-1 forward is
    ~60 launches of a few microseconds each."""

    def __init__(self, model, device=None, bgr=True, mean=IMAGENET_MEAN, std=IMAGENET_STD, use_graphs=True):
        self.model = model.eval()
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.T = model.num_frames
        pe = model.patch_embed
        self.H, self.W = pe.img_size
        self.bgr, self.mean, self.std = bool(bgr), tuple(mean), tuple(std)
        self.frames = torch.zeros(2 * self.T, 3, self.H, self.W, dtype=torch.bfloat16, device=self.device)
        self._stage = torch.empty(1, self.H, self.W, 3, dtype=torch.uint8, device=self.device)
        self._pin = torch.empty(1, self.H, self.W, 3, dtype=torch.uint8).pin_memory()
        self.n = 0
        self.use_graphs = bool(use_graphs)
        self._graphs = {}
        self._graph_token = None
        self._copied = None  # event: the H2D copy out of the pinned staging frame has finished

    def reset(self):
        self.n = 0

    def _forward(self, start):
        return self.model.forward_windows(self.frames, start=start, count=1, stride=1)

    @torch.no_grad()
    def push(self, frame_u8):
        """One new frame (uint8 [H, W, 3], host or device).  Returns None until T frames have been pushed, then the
        (logits, probs) of the current window as host tensors of shape [num_classes]."""
        if frame_u8.dtype != torch.uint8 or frame_u8.dim() != 3 or frame_u8.shape[2] != 3:
            raise ValueError(f"expected a uint8 frame [H, W, 3], got {frame_u8.dtype} {tuple(frame_u8.shape)}")
        if tuple(frame_u8.shape[:2]) != (self.H, self.W):
            # full-size frame: upload it as is and do cv2.resize(..., INTER_CUBIC) (ri:91-92) on the device
            from . import frames as _frames
            raw = frame_u8 if frame_u8.is_cuda else frame_u8.pin_memory().to(self.device, non_blocking=True)
            _frames.resize_cubic_u8(raw[None].contiguous(), (self.H, self.W), out=self._stage)
        elif frame_u8.is_cuda:
            self._stage[0].copy_(frame_u8)
        else:
            if self._copied is not None:
                self._copied.synchronize()  # the previous frame's asynchronous copy still reads the pinned buffer
            self._pin[0].copy_(frame_u8)
            self._stage.copy_(self._pin, non_blocking=True)
            if self._copied is None:
                self._copied = torch.cuda.Event()
            self._copied.record(torch.cuda.current_stream(self.device))
        p = self.n % self.T
        for slot in (p, p + self.T):
            _lib.normalize_frames_u8(self._stage, self.mean, self.std, bgr=self.bgr, out=self.frames[slot:slot + 1])
        self.n += 1
        if self.n < self.T:
            return None
        start = p + 1 if p + 1 < self.T else 0  # oldest frame of the window: slot p + 1 (slot 0 when p = T - 1)
        if not self.use_graphs:
            logits, probs = self._forward(start)
            return logits[0].cpu(), probs[0].cpu()
        # a captured graph has the prepared model's weights and workspace pointers baked in: drop the graphs when either
        # has changed since they were captured (weights reloaded / moved, workspace reallocated by a larger batch elsewhere)
        token = self.model.prepare(self.device).graph_token()
        if token != self._graph_token:
            self._graphs.clear()
        g = self._graphs.get(start)
        if g is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._forward(start)  # warm-up off the capture stream (weight packing, workspace)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._forward(start)
            new_token = self.model.prepare(self.device).graph_token()
            if new_token != self._graph_token:  # (first capture, or the warm-up above moved the workspace)
                self._graphs.clear()
                self._graph_token = new_token
            g = self._graphs[start] = (graph, out)
        g[0].replay()
        return g[1][0][0].cpu(), g[1][1][0].cpu()
