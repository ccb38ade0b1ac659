"""Frame / patch indexing must be bit-exact (BASELINE.json north_star): oracle vs closed form vs C ABI."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import yamnet_oracle as O


def closed_form(n, hop_frames):
    """integer restatement, valid while float32 represents n-15600 exactly (< 2**24)."""
    H = hop_frames * 160
    n2 = max(n, 15600)
    hops = -(-(n2 - 15600) // H)
    N = 15600 + hops * H
    F = 1 + (N - 400) // 160
    P = 1 + (F - 96) // hop_frames
    return N, F, P


@pytest.mark.parametrize("n,hop,expect", [
    (15360, 96, 1), (3194880, 96, 208), (3194880, 48, 415), (57600000, 96, 3750), (57600000, 48, 7499),
    (1382400000, 96, 90000), (0, 96, 1), (1, 96, 1), (15600, 96, 1), (15601, 96, 2), (15601, 48, 2),
])
def test_known_patch_counts(n, hop, expect):
    assert O.frame_counts(n, hop)[2] == expect


@given(st.integers(0, 2 ** 24), st.sampled_from([96, 48]))
@settings(max_examples=300, deadline=None)
def test_oracle_matches_closed_form_below_2p24(n, hop):
    assert O.frame_counts(n, hop) == closed_form(n, hop)


@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([96, 48, 24, 10, 1]))
@settings(max_examples=500, deadline=None)
def test_c_abi_matches_oracle(built_lib, n, hop):
    from buzzdetect_b200 import capi
    assert capi.frames_for(n, hop) == O.frame_counts(n, hop)


def test_float32_ceil_quirk_is_reproduced(built_lib):
    """Above 2**24 samples the float32 quotient can round: the C ABI must follow the graph, not integer ceil."""
    from buzzdetect_b200 import capi
    diffs = 0
    rng = np.random.default_rng(0)
    for n in rng.integers(2 ** 24, 2 ** 31 - 1, size=2000):
        n = int(n)
        a = capi.frames_for(n, 96)
        assert a == O.frame_counts(n, 96)
        diffs += a != closed_form(n, 96)
    assert diffs > 0   # the quirk exists; if this fails the test inputs no longer exercise it


def test_frames_for_rejects_bad_arguments(built_lib):
    from buzzdetect_b200 import capi
    with pytest.raises(ValueError):
        capi.frames_for(-1, 96)
    with pytest.raises(ValueError):
        capi.frames_for(100, 0)
    with pytest.raises(ValueError):
        capi.frames_for(100, 97)


def test_frame_starts_follow_writer_rounding():
    """src/write/formatting.py:5-17."""
    s = O.frame_starts(5, 0.96, 0.0)
    assert s.tolist() == [0.0, 0.96, 1.92, 2.88, 3.84]
    s = O.frame_starts(3, 0.48, 199.68)
    assert s.tolist() == [199.68, 200.16, 200.64]
    from buzzdetect_b200.write import frame_starts
    for hop_s, t0, n in [(0.96, 0.0, 208), (0.48, 199.68, 415), (0.96, 86201.28, 208), (0.48, 0.0, 7499)]:
        assert np.array_equal(frame_starts(n, hop_s, t0, 2), O.frame_starts(n, hop_s, t0, 2))
