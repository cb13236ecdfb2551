"""Window geometry of frame-level inference: the reference's RegularSequencer (dataset/sequencing.py:32-62) as a
closed-form plan instead of lists of frame indices.

A video of `timesteps_nb` frames recorded at `input_frequency` fps is scored at `seq_frequency` fps with windows of
`seq_length` frames every `step` source frames (dota.py:209-213: 10 fps source, 16 frames, step 1; dada.py:31,173-177:
30 fps source scored at 10 fps, i.e. every third frame inside a window).  The windows are aligned to the END of the
video: the last window always ends on the last frame (sequencing.py:53-60).

    plan = window_plan(T, input_frequency=30, seq_frequency=10, seq_length=16, step=1)
    model.forward_windows(frames, start=plan.start, count=plan.count, stride=plan.stride, frame_step=plan.frame_step)

`WindowPlan.sequences()` expands the plan into the index lists RegularSequencer.get_sequences returns (tests compare
them with the unmodified reference class).
"""
from collections import namedtuple


class WindowPlan(namedtuple("WindowPlan", "start count stride frame_step length")):
    """frame t of window w is  start + w * stride + t * frame_step,  0 <= w < count, 0 <= t < length."""
    __slots__ = ()

    @property
    def span(self):
        """Source frames one window covers (`actual_seq_length`, sequencing.py:50)."""
        return (self.length - 1) * self.frame_step + 1

    def last_frames(self):
        """Index of the last frame of every window — the frame a window's label / score belongs to (dota.py:217-223)."""
        return [self.start + w * self.stride + self.span - 1 for w in range(self.count)]

    def sequences(self):
        return [list(range(s, s + self.length * self.frame_step, self.frame_step))
                for s in range(self.start, self.start + self.count * self.stride, self.stride)]


def window_plan(timesteps_nb, input_frequency=10, seq_frequency=10, seq_length=16, step=1):
    """RegularSequencer(seq_frequency, seq_length, step).get_sequences(timesteps_nb, input_frequency) as a WindowPlan;
    None when the video is shorter than one window (sequencing.py:51-52)."""
    if hasattr(timesteps_nb, "__len__"):
        timesteps_nb = len(timesteps_nb)
    if seq_frequency <= 0 or input_frequency <= 0 or step <= 0:
        raise ValueError("frequencies and step must be positive")
    if input_frequency % seq_frequency != 0:
        raise ValueError(f"input frequency {input_frequency} must be divisible by the target frequency {seq_frequency}")
    if isinstance(seq_length, float):                       # seconds -> frames (sequencing.py:11-14)
        seq_length = round(seq_length * seq_frequency)
    frame_step = input_frequency // seq_frequency
    span = (seq_length - 1) * frame_step + 1
    if span > timesteps_nb:
        return None
    count = (timesteps_nb - span) // step + 1
    start = (timesteps_nb - span) % step
    return WindowPlan(start, count, step, frame_step, seq_length)


def window_plan_with_start(timesteps_nb, input_frequency=10, seq_frequency=10, seq_length=16, step=1):
    """RegularSequencerWithStart (dataset/sequencing.py:132-167; the DAPT video datasets, dota.py:555): the plan of
    window_plan plus, when end alignment leaves more than min(0.3 * input_frequency, 5) frames unused at the start of the
    video, ONE extra window that begins at frame 0 (the reference appends it after the regular ones).
    Returns (plan, extra) with extra = WindowPlan(start=0, count=1, ...) or None; (None, None) for a too-short video."""
    plan = window_plan(timesteps_nb, input_frequency, seq_frequency, seq_length, step)
    if plan is None:
        return None, None
    extra = None
    if plan.start > min(0.3 * input_frequency, 5):
        extra = WindowPlan(0, 1, plan.stride, plan.frame_step, plan.length)
    return plan, extra
