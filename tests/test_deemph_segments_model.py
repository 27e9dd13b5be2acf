"""The algorithm of the segmented de-emphasis (csrc/kernels.cuh k_deemph / k_deemph_fix), restated in numpy and checked
against the plain serial chain (ref src/celt.cpp:1977-1982: y = x + m; m = (27853 * y) >> 15) -- CPU only.

The GPU kernels are tested bit-exact against the reference decoder in tests/test_gpu_celt.py::test_deemphasis_segments; this
file pins the ARGUMENT they rest on: a run from memory 0 that warms up on the samples in front of a segment either arrives
at the serial run's memory -- then everything behind is identical -- or is caught by the comparison and redone, and a redo
may stop at the first frame end where it meets the segment's own run."""
import numpy as np

N = 960


def serial(x, m0):
    """-> (y[n], memory after every sample)"""
    m = int(m0)
    y, ms = [], []
    for v in x.tolist():
        t = v + m
        m = (27853 * t) >> 15
        y.append(t)
        ms.append(m)
    return np.array(y, np.int64), np.array(ms, np.int64)


def segmented(x, m0, n_frames, segs, warm):
    """The two kernels: returns (y, final memory, segments redone, frames redone)."""
    y = np.zeros(n_frames * N, np.int64)
    bounds = [n_frames * g // segs for g in range(segs + 1)]
    spec, end, memf = {}, {}, {}
    for g in range(segs):  # k_deemph: every segment on its own
        a, b = bounds[g] * N, bounds[g + 1] * N
        m = m0 if g == 0 else 0
        if g:
            _, ms = serial(x[a - warm:a], 0)
            m = int(ms[-1])
        spec[g] = m
        yy, ms = serial(x[a:b], m)
        y[a:b] = yy
        end[g] = int(ms[-1])
        for f in range(bounds[g], bounds[g + 1]):
            memf[f] = int(ms[(f - bounds[g] + 1) * N - 1])
    m = end[0]
    redone = frames = 0
    for g in range(1, segs):  # k_deemph_fix: the segments in order
        if spec[g] == m:
            m = end[g]
            continue
        redone += 1
        for f in range(bounds[g], bounds[g + 1]):
            yy, ms = serial(x[f * N:(f + 1) * N], m)
            y[f * N:(f + 1) * N] = yy
            m = int(ms[-1])
            frames += 1
            if m == memf[f]:
                m = end[g]
                break
    return y, m, redone, frames


def signal(rng, n_frames, kind):
    x = (rng.standard_normal(n_frames * N) * 10 ** rng.uniform(0.5, 7.5)).astype(np.int64)
    if kind == "silent_frames":  # digital silence in whole frames (the silence flag): the case that makes a warm-up miss
        for f in rng.choice(n_frames, n_frames // 3, replace=False):
            x[f * N:(f + 1) * N] = 0
    elif kind == "constant":
        x[:] = int(rng.integers(-5, 6))
    return x


def test_segmented_run_equals_the_serial_chain():
    rng = np.random.default_rng(7)
    redo_total = 0
    for trial in range(60):
        kind = ("noise", "silent_frames", "constant")[trial % 3]
        n_frames = int(rng.integers(4, 13))
        segs = int(rng.choice([2, 3, 4]))
        segs = min(segs, n_frames // 2)
        warm = int(rng.choice([32, 96, 256]))
        x = signal(rng, n_frames, kind)
        m0 = int(rng.integers(-2 ** 27, 2 ** 27))
        want, ms = serial(x, m0)
        got, m_end, redone, frames = segmented(x, m0, n_frames, segs, warm)
        assert np.array_equal(got, want), (trial, kind, segs, warm)
        assert m_end == int(ms[-1])
        redo_total += redone
    assert redo_total > 0  # short warm-ups and silent frames: the redo path carried some of these


def test_full_warm_up_meets_the_serial_run_on_signal_and_misses_on_silence():
    rng = np.random.default_rng(8)
    # noise of any loudness: 256 samples of warm-up always arrive at the serial run's memory
    for _ in range(40):
        x = signal(rng, 4, "noise")
        _, _, redone, _ = segmented(x, int(rng.integers(-2 ** 27, 2 ** 27)), 4, 2, 256)
        assert redone == 0
    # a silent frame in front of the second segment and a negative memory: on zeros the serial run settles at a negative
    # fixed point of m = floor(0.85 m) (-6 .. -1), the warm-up from 0 stays at 0;
    # the redo meets the segment's own run within its first frame
    x = signal(rng, 6, "noise")
    x[2 * N:3 * N] = 0
    want, ms = serial(x, -12345)
    got, m_end, redone, frames = segmented(x, -12345, 6, 2, 256)
    assert -6 <= int(ms[3 * N - 1]) <= -1
    assert redone == 1 and frames == 1
    assert np.array_equal(got, want) and m_end == int(ms[-1])
