"""CPU model check of the attention kernel's mbarrier protocol (tools/attn_protocol_check.py): every interleaving of the
issuing warps, the in-order tensor pipe and the softmax warps for small tile counts.

Parity waits only tell consecutive phases apart, so a scheme can be correct for well-behaved timing and still corrupt
data or deadlock when warps drift apart (both were seen on the GPU with an experimental softmax at the end of round 1).
These tests pin down (1) that the SAFE scheme (attn_fwd2_kernel<..., SAFE=true>: barriers per (buffer, tile) pair, o_done
awaited every step) needs no timing assumption, (2) the exact assumptions the default scheme relies on, and (3) that the
checker is not vacuous: it finds the premature pass as soon as one of those assumptions is dropped."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from attn_protocol_check import check  # noqa: E402


@pytest.mark.parametrize("warps", [1, 2])
@pytest.mark.parametrize("n_tiles", [1, 2, 3, 4, 5])
def test_safe_scheme_needs_no_timing_assumption(n_tiles, warps):
    ok, msg, states = check(n_tiles, warps, "v5")
    assert ok, msg
    assert states > 50


@pytest.mark.parametrize("n_tiles", [1, 2, 3, 4, 5, 6])
def test_default_scheme_is_safe_under_its_two_timing_assumptions(n_tiles):
    ok, msg, _ = check(n_tiles, 2, "v2", mma_keeps_up=True, slots_in_order=True)
    assert ok, msg


def test_default_scheme_without_mma_keeping_up_passes_o_done_early():
    ok, msg, _ = check(2, 2, "v2", mma_keeps_up=False, slots_in_order=True)
    assert not ok and "premature pass" in msg and "o_done" in msg


def test_default_scheme_without_slot_order_passes_p_full_early():
    ok, msg, _ = check(2, 2, "v2", mma_keeps_up=True, slots_in_order=False)
    assert not ok and "premature pass" in msg and "issuer" in msg


def test_each_ingredient_removes_one_assumption():
    for n_tiles in (2, 3, 4):
        ok, msg, _ = check(n_tiles, 2, "six_barriers_only", mma_keeps_up=True, slots_in_order=False)
        assert ok, msg
        ok, msg, _ = check(n_tiles, 2, "o_done_always_only", mma_keeps_up=False, slots_in_order=True)
        assert ok, msg
    assert not check(2, 2, "six_barriers_only", mma_keeps_up=False)[0]
    assert not check(2, 2, "o_done_always_only", slots_in_order=False)[0]


# ----------------------------------------------------------------------------------------------------------------------
# v3 / v4 (double-buffered score tiles per query tile): tools/attn_protocol_check_v4.py
def _check_v4(*a, **k):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "attn_protocol_check_v4", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                                               "attn_protocol_check_v4.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.check(*a, **k)


def test_v4_protocol_is_safe_for_every_interleaving():
    """s_full / p_full per score buffer, o_done awaited only by the online form's rescale, the epilogue on o_final: no
    premature pass, no hazard, no deadlock for any interleaving, with and without the rescale path, 2 and 3 softmax warps."""
    for n_tiles in range(1, 7):
        for warps in (2, 3):
            for online in (False, True):
                ok, msg, _ = _check_v4(n_tiles, warps, "o_final", online)
                assert ok, (n_tiles, warps, online, msg)


def test_v3_first_epilogue_wait_was_a_premature_pass():
    """What v3 first shipped (epilogue waits on o_done with the parity of the last phase): a warp may get there while P V of
    the step before is still outstanding, the parity wait passes two phases early and O is read too soon - the model check
    finds it for every tile count >= 2 (on the GPU: 3 of 36 split-KV runs off by 5e-3..8e-3)."""
    assert _check_v4(1, 2, "o_done_parity")[0]
    for n_tiles in range(2, 6):
        ok, msg, _ = _check_v4(n_tiles, 2, "o_done_parity")
        assert not ok and "premature pass" in msg and "epilogue" in msg
