"""Integer contracts: the oracle restatement is pinned on the survey's known answers
and the product's host logic must agree with it bit-for-bit."""
import random

import pytest
import torch
from hypothesis import given, settings, strategies as st

from affectgpt_b200 import sampling as S
from oracle import sampling as O


def test_uniform_known_answers():
    # A/demo/sample_00000000.mp4 has 88 frames (SURVEY 8c fixtures)
    assert O.uniform_indices(88) == [0, 11, 22, 33, 44, 55, 66, 77]
    assert O.uniform_indices(37) == [0, 4, 9, 13, 18, 23, 27, 32]
    assert O.uniform_indices(5) == [0, 1, 2, 3, 4, 4, 4, 4]
    assert O.uniform_indices(1) == [0] * 8


@given(st.integers(min_value=1, max_value=200000), st.sampled_from([1, 4, 8, 16, 32]))
@settings(max_examples=400, deadline=None)
def test_uniform_matches_oracle(vlen, n):
    got = S.uniform_frame_indices(vlen, n)
    assert got == O.uniform_indices(vlen, n)
    assert len(got) == n and all(0 <= i < vlen for i in got) and got == sorted(got)


def test_audio_known_answers():
    # A/demo/sample_00000000.wav has 55357 samples
    exp = {55357: [0, 3336, 6673, 10010, 13346, 16683, 20020, 23356],
           64000: [0, 4571, 9142, 13714, 18285, 22857, 27428, 32000],
           128000: [0, 13714, 27428, 41142, 54857, 68571, 82285, 96000],
           32000: [0] * 8, 100: [0] * 8}
    for n, starts in exp.items():
        _, b = O.audio_segments(torch.zeros(1, n))
        assert [x[0] for x in b] == starts
        assert S.audio_segment_bounds(n) == b


@given(st.integers(min_value=1, max_value=16000 * 120))
@settings(max_examples=300, deadline=None)
def test_audio_bounds_match_oracle(n):
    _, b = O.audio_segments(torch.zeros(1, n))
    assert S.audio_segment_bounds(n) == b
    assert len(b) == 8 and all(e - s == 32000 for s, e in b)


def test_segment_audio_matches_oracle_including_stereo_and_short():
    g = torch.Generator().manual_seed(0)
    for ch, n in ((1, 55357), (2, 70001), (1, 1234), (2, 31999), (1, 32000)):
        w = torch.randn(ch, n, generator=g)
        a, _ = O.audio_segments(w)
        assert torch.equal(S.segment_audio(w), a)
        assert a.shape == (8, 1, 32000)


def test_emotion_peak_fallbacks():
    assert O.emotion_peak_indices(88, None) == [44] * 8          # no au_info -> middle frame repeated
    assert S.emotion_peak_frame_indices(88, None) == [44] * 8
    for v in range(1, 200):
        assert S.smart_frame_indices(None, v) == O.smart_frame_indices(None, v)


def test_smart_indices_match_oracle_randomised():
    rng = random.Random(1)
    for _ in range(4000):
        total = rng.randint(1, 80)
        p = rng.randint(0, total - 1)
        au = {"peak_frames": [{"peak_index": p, "frames_before_peak": rng.randint(0, 4),
                               "frames_after_peak": rng.randint(0, 4)}], "total_frames": total}
        got = S.smart_frame_indices(au, total)
        assert got == O.smart_frame_indices(au, total)
        assert len(got) == 8 or total < 8 and len(got) == 8
        assert S.emotion_peak_frame_indices(total, au) == O.emotion_peak_indices(total, au)


def test_smart_indices_example():
    au = {"peak_frames": [{"peak_index": 40, "frames_before_peak": 3, "frames_after_peak": 3}], "total_frames": 88}
    got = O.smart_frame_indices(au, 88)
    assert {38, 39, 40, 41, 42} <= set(got) and len(got) == 8 and got == sorted(got)


def test_demo_clip_fixture_follows_the_reference_rules():
    """tests/golden/golden_demo_clip.npz (made from A/demo/sample_00000000.{mp4,wav}): 8 frames at the uniform indices of an
    88-frame video, 55 357 PCM samples whose eight 2-second slices start where SURVEY 8a-a4 says."""
    import os
    import numpy as np
    from affectgpt_b200 import sampling
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_demo_clip.npz"))
    assert gold["frames_u8"].shape == (8, 224, 224, 3) and gold["frames_u8"].dtype == np.uint8
    assert gold["pcm_i16"].shape == (55357,)
    assert sampling.uniform_frame_indices(88, 8) == [0, 11, 22, 33, 44, 55, 66, 77]
    assert [a for a, _ in sampling.audio_segment_bounds(55357, 16000, 2, 8)] == [0, 3336, 6673, 10010, 13346, 16683, 20020, 23356]
    assert gold["frame_hidden"].shape == (1, 8, 768) and gold["audio_hidden"].shape == (1, 8, 1024)
    assert gold["token_multi"].shape == (1, 1, 3584) and np.isfinite(gold["token_multi"]).all()
