"""CPU emulation of the fixed-reference softmax of attn_fwd2_kernel<..., FIXEDREF=true> (experimental instance).

With QK-RMSNorm (transformer_sd3.py: qk_norm="rms_norm"; attention.py:484-489,502-505) every row of q and k has
norm <= 8 * max|w| (RMSNorm over head_dim 64 makes the unweighted norm exactly sqrt(64)), so
    score * scale * log2(e) <= (8 max|w_q|) (8 max|w_k|) / 8 * log2(e) =: ref
and p = 2^(s*scale*log2e - ref) can be taken against that FIXED reference: no online row max, no rescaling of the
accumulator.  Softmax is invariant to the reference; what has to be shown is that the kernel's arithmetic (fp32
exponentials and row sums, P rounded to bf16 for the tensor core, fp32 accumulation) keeps its accuracy for any
admissible reference, and where it stops working (underflow) - the host must fall back to the online max there."""
import math

import torch

LOG2E = 1.4426950408889634


def rms_norm(x, w, eps=1e-6):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def emulate(q, k, v, ref_log2, scale=0.125):
    """q, k, v: bf16 [S, 64] -> what the kernel computes with a fixed reference (fp32 math, bf16 P)."""
    s = q.float() @ k.float().T                                      # tensor core: bf16 operands, fp32 accumulate
    p = torch.exp2(s * (scale * LOG2E) - ref_log2)                   # fp32
    l = p.sum(-1, keepdim=True)                                      # row sum of the fp32 values (as in the kernel)
    o = p.bfloat16().float() @ v.float()                             # P rounded to bf16 feeds the P V MMA
    return o / l


def exact(q, k, v, scale=0.125):
    s = (q.double() @ k.double().T) * scale
    return torch.softmax(s, -1) @ v.double()


def make(seed, S=700, wscale=0.15):
    g = torch.Generator().manual_seed(seed)
    wq = 1 + wscale * torch.randn(64, generator=g)
    wk = 1 + wscale * torch.randn(64, generator=g)
    q = rms_norm(torch.randn(S, 64, generator=g) * 3, wq).bfloat16()
    k = rms_norm(torch.randn(S, 64, generator=g) * 0.2, wk).bfloat16()
    v = torch.randn(S, 64, generator=g).bfloat16()
    bound = (8 * wq.abs().max()) * (8 * wk.abs().max()) / 8 * LOG2E
    return q, k, v, float(bound)


def rel(a, b):
    return ((a.double() - b).norm() / b.norm()).item()


def test_bound_holds_for_rms_normed_q_k():
    for seed in range(5):
        q, k, v, bound = make(seed)
        smax = ((q.float() @ k.float().T) * 0.125 * LOG2E).max().item()
        assert smax <= bound * 1.01          # bf16 rounding of q, k can add a fraction of a percent
        assert bound < 40


def test_fixed_reference_matches_exact_softmax_for_any_admissible_reference():
    q, k, v, bound = make(0)
    ref = exact(q, k, v)
    errs = [rel(emulate(q, k, v, r), ref) for r in (bound, bound + 10, 40.0, 60.0)]
    assert max(errs) < 3e-3, errs                      # the bf16 rounding of P, as with the online max
    assert max(errs) - min(errs) < 5e-4, errs          # independent of the reference


def test_fixed_reference_breaks_down_by_underflow_far_above_the_bound():
    """Why the host must only choose this mode for small bounds: at ref ~ 150 every p flushes to zero."""
    q, k, v, bound = make(0)
    o = emulate(q, k, v, 160.0)
    assert not torch.isfinite(o).all()


def test_peaked_rows_keep_their_accuracy():
    """Aligned q and k (scores close to the bound) next to rows whose scores are all far below it."""
    g = torch.Generator().manual_seed(3)
    w = torch.ones(64)
    base = torch.randn(1, 64, generator=g)
    q = rms_norm(torch.cat([base.repeat(8, 1), torch.randn(120, 64, generator=g)]), w).bfloat16()
    k = rms_norm(torch.cat([base.repeat(4, 1), -base.repeat(4, 1), torch.randn(248, 64, generator=g)]), w).bfloat16()
    v = torch.randn(256, 64, generator=g).bfloat16()
    bound = 8 * 8 / 8 * LOG2E
    assert rel(emulate(q, k, v, bound), exact(q, k, v)) < 3e-3


def test_prescaled_q_norm_weights_fold_scale_and_bound_the_scores():
    """engine.prescaled_q_norm_weights (DIT4SR_ATTN_XREG=4): (x_q * w_q') . k == log2(e)/sqrt(64) * (x_q * w_q) . k, the
    returned bound dominates every score of RMS-normed rows (image and text streams), and large weights are refused."""
    import pytest
    from dit4sr_b200.engine import prescaled_q_norm_weights
    g = torch.Generator().manual_seed(0)
    wq, wk = 1 + 0.2 * torch.randn(64, generator=g), 1 + 0.2 * torch.randn(64, generator=g)
    waq, wak = 1 + 0.2 * torch.randn(64, generator=g), 1 + 0.2 * torch.randn(64, generator=g)
    wq2, waq2, bound = prescaled_q_norm_weights(wq, wk, waq, wak)
    xq, xk = torch.randn(300, 64, generator=g) * 5, torch.randn(300, 64, generator=g) * 0.3
    one = torch.ones(64)
    for a, a2 in ((wq, wq2), (waq, waq2)):
        for b in (wk, wak):
            k = rms_norm(xk, b)
            kk = rms_norm(k, b)                                    # the twice-normed LR keys (quirk Q1)
            for keys in (k, kk):
                s_ref = (rms_norm(xq, a) @ keys.T) * (0.125 * LOG2E)
                s_new = rms_norm(xq, one) * a2 @ keys.T
                assert torch.allclose(s_ref, s_new, rtol=1e-5, atol=1e-5)
                assert s_new.abs().max().item() <= bound
    with pytest.raises(ValueError):
        prescaled_q_norm_weights(wq * 3, wk * 3)


def test_prescaled_scores_without_any_reference_keep_the_accuracy():
    """DIT4SR_ATTN_XREG=4: q arrives as bf16(q * log2(e)/sqrt(64)) (the scale is folded into the q RMSNorm weights before
    the projection epilogue rounds to bf16) and the kernel takes p = 2^score with no reference.  Against the exact softmax
    of the unscaled fp32 q the only new error is that one extra bf16 rounding of q."""
    g = torch.Generator().manual_seed(5)
    wq = 1 + 0.15 * torch.randn(64, generator=g)
    wk = 1 + 0.15 * torch.randn(64, generator=g)
    xq, xk = torch.randn(600, 64, generator=g) * 2, torch.randn(600, 64, generator=g)
    v = torch.randn(600, 64, generator=g).bfloat16()
    f = LOG2E / 8.0
    q32, k = rms_norm(xq, wq), rms_norm(xk, wk).bfloat16()
    q_scaled = (rms_norm(xq, torch.ones(64)) * (wq * f)).bfloat16()      # what the QKV_NORM epilogue would store
    o = emulate(q_scaled, k, v, ref_log2=0.0, scale=1.0 / LOG2E)         # p = 2^(q' . k)
    assert torch.isfinite(o).all()
    ref = exact(q32, k, v)                                               # exact softmax of the unrounded q
    base = emulate(q32.bfloat16(), k, v, ref_log2=20.0)                  # today's arithmetic (q rounded unscaled)
    assert rel(o, ref) < 4e-3
    assert rel(o, ref) < 1.5 * rel(base, ref) + 1e-3
