"""Window geometry (dataset/sequencing.py:32-62): simple_tad_b200.sequencing.window_plan against the index lists the
unmodified RegularSequencer produced (tests/golden/sequencer.npz, oracle/make_golden.py), and the runner's own window
arithmetic against the plan.  CPU only."""
import numpy as np
import pytest

from simple_tad_b200 import sequencing
from simple_tad_b200.runner import window_segments
from tests import parity


def test_window_plan_reproduces_regular_sequencer():
    g = parity.golden("sequencer")
    for i, (T, fin, fseq, length, step) in enumerate(g["cases"].tolist()):
        ref = g[f"seq_{i}"]
        plan = sequencing.window_plan(T, fin, fseq, length, step)
        if ref.shape[0] == 0:
            assert plan is None
            continue
        seqs = np.array(plan.sequences())
        assert seqs.shape == ref.shape and np.array_equal(seqs, ref), (T, fin, fseq, length, step)
        assert plan.last_frames() == ref[:, -1].tolist() and ref[-1, -1] == T - 1      # aligned to the last frame
        assert plan.frame_step == fin // fseq and plan.span == (length - 1) * plan.frame_step + 1
        segs, total = window_segments([T], 0, 10 ** 9, length, step, plan.frame_step)
        assert total == plan.count and segs == [(0, 0, plan.count)]


def test_window_plan_with_start_reproduces_the_reference():
    """RegularSequencerWithStart (dataset/sequencing.py:132-167): the regular windows, then one window from frame 0 when
    the end-aligned windows leave the first frames uncovered."""
    g = parity.golden("sequencer")
    extras = 0
    for i, (T, fin, fseq, length, step) in enumerate(g["cases"].tolist()):
        ref = g[f"seqws_{i}"]
        plan, extra = sequencing.window_plan_with_start(T, fin, fseq, length, step)
        if ref.shape[0] == 0:
            assert plan is None and extra is None
            continue
        seqs = plan.sequences() + (extra.sequences() if extra is not None else [])
        assert np.array_equal(np.array(seqs), ref), (T, fin, fseq, length, step)
        extras += extra is not None
    assert extras >= 2   # the grid holds cases on both sides of the min(0.3 * fps, 5) rule


def test_window_plan_arguments():
    assert sequencing.window_plan(range(100)).count == 85                 # a sequence of timesteps (sequencing.py:43-44)
    assert sequencing.window_plan(100, seq_length=1.6).length == 16       # seconds (sequencing.py:11-14)
    with pytest.raises(ValueError):
        sequencing.window_plan(100, input_frequency=25, seq_frequency=10)
    with pytest.raises(ValueError):
        sequencing.window_plan(100, step=0)
