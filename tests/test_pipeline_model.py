"""The planning model of the forward pipeline (tools/pipeline_model.py) must keep reproducing what was measured for
the shipped schedule - otherwise its prediction for the next schedule means nothing."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    spec = importlib.util.spec_from_file_location("pipeline_model", os.path.join(ROOT, "tools", "pipeline_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_shipped_schedule_matches_the_measured_timeline():
    m = _model()
    period, busy = m.two_stage()
    # profiles/timeline_r01_full_4096_1thread_per_row.txt: 2760-2930 clocks per iteration; ncu / bench: 0.66-0.73
    assert 2750 <= period <= 2950
    assert 0.68 <= busy <= 0.75


def test_pipelined_schedule_needs_three_s_buffers():
    m = _model()
    assert m.pipelined(s_buffers=2)[1] < 0.75       # two buffers leave the chain as long as it is today
    assert m.pipelined(s_buffers=3)[1] > 0.95
    slow = m.Lat(rowmax=int(413 * 1.3), exp34=int(940 * 1.3), exp_last=int(383 * 1.3))
    assert m.pipelined(lat=slow)[1] > 0.8           # still ahead with a 30 % slower softmax
