"""GPU parity tests of each kernel, called through the C-ABI, against plain torch fp32 math on the same inputs.

Tolerances: outputs stored as bf16 carry one rounding (2^-9 relative) -> rel-L2 <= 3e-3; fp32 outputs <= 2e-5.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_TOL = 3e-3
F32_TOL = 2e-5


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from dit4sr_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("B,R,D", [(2, 333, 1536), (4, 2048, 1536), (1, 7, 128), (3, 1, 1536)])
def test_ln_modulate(ops, B, R, D):
    torch.manual_seed(0)
    x = torch.randn(B, R, D, device="cuda") * 2 + 0.5
    mod = torch.randn(B, 4 * D, device="cuda") * 0.5
    sh, sc, sh2, sc2 = mod[:, :D], mod[:, D:2 * D], mod[:, 2 * D:3 * D], mod[:, 3 * D:]
    o1 = torch.empty(B, R, D, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    ops.ln_modulate(x, sh, sc, o1, sh2, sc2, o2)
    ln = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
    assert rel(o1.float(), ln * (1 + sc[:, None]) + sh[:, None]) < BF16_TOL
    assert rel(o2.float(), ln * (1 + sc2[:, None]) + sh2[:, None]) < BF16_TOL


@pytest.mark.parametrize("B,R,Rb,D,dual,f32", [(2, 2048, 333, 1536, False, False), (3, 700, 333, 1536, True, False),
                                               (1, 5, 9, 128, False, False), (8, 1, 1, 1536, True, False),
                                               (2, 300, 77, 1536, False, True)])
def test_ln_modulate_companion(ops, B, R, Rb, D, dual, f32):
    """ABI v5: the text stream's rows normalised by the image stream's launch (own modulation vectors, single output);
    bit-identical to two separate launches."""
    torch.manual_seed(0)
    x = torch.randn(B, R, D, device="cuda") * 2 + 0.5
    xb = torch.randn(B, Rb, D, device="cuda") * 3 - 0.25
    mod = torch.randn(B, 6 * D, device="cuda") * 0.5
    sh, sc, sh2, sc2, shb, scb = (mod[:, i * D:(i + 1) * D] for i in range(6))
    dt = torch.float32 if f32 else torch.bfloat16
    o1, o2, ob = (torch.full((B, r, D), float("nan"), device="cuda", dtype=dt) for r in (R, R, Rb))
    if dual:
        ops.ln_modulate(x, sh, sc, o1, sh2, sc2, o2, companion=(xb, shb, scb, ob))
    else:
        ops.ln_modulate(x, sh, sc, o1, companion=(xb, shb, scb, ob))
    tol = F32_TOL if f32 else BF16_TOL
    ln, lnb = (torch.nn.functional.layer_norm(t, (D,), eps=1e-6) for t in (x, xb))
    assert rel(o1.float(), ln * (1 + sc[:, None]) + sh[:, None]) < tol
    assert rel(ob.float(), lnb * (1 + scb[:, None]) + shb[:, None]) < tol
    if dual:
        assert rel(o2.float(), ln * (1 + sc2[:, None]) + sh2[:, None]) < tol
    s1, sb = torch.empty_like(o1), torch.empty_like(ob)
    ops.ln_modulate(x, sh, sc, s1)
    ops.ln_modulate(xb, shb, scb, sb)
    assert torch.equal(s1, o1) and torch.equal(sb, ob)


@pytest.mark.parametrize("B,G,C", [(2, 32, 6144), (1, 5, 512), (1, 64, 256), (1, 1, 8)])
def test_dwconv_matches_conv2d(ops, B, G, C):
    torch.manual_seed(0)
    N = G * G
    u = torch.randn(B, 2 * N, C, device="cuda").bfloat16()
    w = torch.randn(C, 1, 3, 3, device="cuda") * 0.1
    bias = torch.randn(C, device="cuda") * 0.1
    lr = u[:, N:].float().reshape(B, G, G, C).permute(0, 3, 1, 2)
    conv = torch.nn.functional.conv2d(lr, w, bias, padding=1, groups=C).reshape(B, C, N).permute(0, 2, 1)
    ref = u[:, :N].float() + conv
    u2 = u.clone()
    ops.dwconv3x3_add(u2[:, :N], u2[:, N:], w.reshape(C, 9).t().contiguous(), bias, G)
    assert rel(u2[:, :N].float(), ref) < BF16_TOL
    assert torch.equal(u2[:, N:], u[:, N:])


def _mk(B, R, K, N):
    A = torch.randn(B, R, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda") * 0.1
    return A, W, bias, A.float() @ W.float().t() + bias


@pytest.mark.parametrize("B,R,K,N", [(1, 128, 64, 256), (1, 256, 1536, 1536), (2, 333, 1536, 4608), (1, 8, 1536, 512),
                                      (1, 1024, 1536, 64), (1, 2048, 6144, 1536), (1, 1, 96, 32), (3, 129, 80, 160)])
def test_gemm_bias(ops, B, R, K, N):
    torch.manual_seed(0)
    A, W, bias, ref = _mk(B, R, K, N)
    o = torch.empty(B, R, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, o, ops.EPI_BIAS)
    o32 = torch.empty(B, R, N, device="cuda")
    ops.gemm(A, W, bias, o32, ops.EPI_BIAS)
    assert rel(o.float(), ref) < BF16_TOL
    assert rel(o32, ref) < F32_TOL


def test_gemm_linearity_full_size(ops):
    """Size-independent property at the full FF1 shape: gemm(A1 + A2) == gemm(A1) + gemm(A2) - bias (fp32 out)."""
    torch.manual_seed(0)
    R, K, N = 4096, 1536, 6144
    A1 = torch.randn(R, K, device="cuda").bfloat16()
    A2 = (A1.float() * 2).bfloat16()       # exactly representable sum: A1 + A2 = 3*A1 needs care -> use power of two
    A3 = (A1.float() * 4).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    z = torch.zeros(N, device="cuda")
    o1, o3 = torch.empty(R, N, device="cuda"), torch.empty(R, N, device="cuda")
    ops.gemm(A1, W, z, o1, ops.EPI_BIAS)
    ops.gemm(A3, W, z, o3, ops.EPI_BIAS)
    assert rel(o3, 4 * o1) < 1e-6         # scaling by a power of two is exact in every partial product
    assert A2 is not None


def test_gemm_strided_views(ops):
    torch.manual_seed(0)
    A, W, bias, ref = _mk(2, 200, 512, 256)
    big = torch.zeros(2, 200, 1024, device="cuda", dtype=torch.bfloat16)
    big[:, :, 512:] = A
    cbuf = torch.zeros(2, 500, 768, device="cuda", dtype=torch.bfloat16)
    ops.gemm(big[:, :, 512:], W, bias, cbuf[:, 100:300, 256:512], ops.EPI_BIAS)
    assert rel(cbuf[:, 100:300, 256:512].float(), ref) < BF16_TOL
    cbuf[:, 100:300, 256:512] = 0
    assert cbuf.abs().max() == 0


def test_gemm_activations(ops):
    torch.manual_seed(0)
    A, W, bias, ref = _mk(1, 512, 1536, 6144)
    o = torch.empty(1, 512, 6144, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, o, ops.EPI_BIAS_GELU)
    assert rel(o.float(), torch.nn.functional.gelu(ref, approximate="tanh")) < BF16_TOL
    ops.gemm(A, W, bias, o, ops.EPI_BIAS_SILU)
    assert rel(o.float(), torch.nn.functional.silu(ref)) < BF16_TOL


def test_gemm_gate_residual(ops):
    torch.manual_seed(0)
    A, W, bias, ref = _mk(2, 333, 1536, 1536)
    gate = torch.randn(2, 1536, device="cuda")
    x = torch.randn(2, 333, 1536, device="cuda")
    x0 = x.clone()
    ops.gemm(A, W, bias, x, ops.EPI_GATE_RESID, gate=gate)
    assert rel(x, x0 + gate[:, None] * ref) < F32_TOL
    pos = torch.randn(1, 333, 1536, device="cuda")
    y = torch.empty(2, 333, 1536, device="cuda")
    ops.gemm(A, W, bias, y, ops.EPI_GATE_RESID, resid=pos)
    assert rel(y, pos + ref) < F32_TOL


def test_gemm_grouped_launch(ops):
    """Three independent problems (different M, N, outputs) in one persistent launch == three single launches."""
    torch.manual_seed(0)
    probs, refs, outs = [], [], []
    for (B, R, K, N) in [(2, 256, 1536, 1536), (2, 333, 1536, 512), (1, 40, 256, 4608)]:
        A, W, bias, ref = _mk(B, R, K, N)
        o = torch.zeros(B, R, N, device="cuda", dtype=torch.bfloat16)
        probs.append(dict(A=A, W=W, bias=bias, out=o, epilogue=ops.EPI_BIAS_GELU))
        refs.append(torch.nn.functional.gelu(ref, approximate="tanh"))
        outs.append(o)
    ops.gemm_grouped(probs)
    for o, r in zip(outs, refs):
        assert rel(o.float(), r) < BF16_TOL
    # mixed gate/residual group: image rows in place + text rows with another gate
    A1, W1, b1, r1 = _mk(2, 300, 512, 256)
    A2, W2, b2, r2 = _mk(2, 77, 512, 256)
    x1, x2 = torch.randn(2, 300, 256, device="cuda"), torch.randn(2, 77, 256, device="cuda")
    g1, g2 = torch.randn(2, 256, device="cuda"), torch.randn(2, 256, device="cuda")
    e1, e2 = x1 + g1[:, None] * r1, x2 + g2[:, None] * r2
    ops.gemm_grouped([dict(A=A1, W=W1, bias=b1, out=x1, epilogue=ops.EPI_GATE_RESID, gate=g1),
                      dict(A=A2, W=W2, bias=b2, out=x2, epilogue=ops.EPI_GATE_RESID, gate=g2)])
    assert rel(x1, e1) < F32_TOL and rel(x2, e2) < F32_TOL


@pytest.mark.parametrize("rows,K,N", [(10240, 256, 512), (6400, 128, 768), (25344 - 100, 64, 256)])
def test_gemm_tail_split(ops, rows, K, N):
    """Tile counts just over a multiple of the SM count: the last tiles are cut into 64- or 128-column slices (tail
    splitting, gemm.cu); every epilogue must give the same result as without it."""
    torch.manual_seed(0)
    A, W, bias, ref = _mk(1, rows, K, N)
    o = torch.zeros(1, rows, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, o, ops.EPI_BIAS_GELU)
    assert rel(o.float(), torch.nn.functional.gelu(ref, approximate="tanh")) < BF16_TOL
    gate = torch.randn(1, N, device="cuda")
    x = torch.randn(1, rows, N, device="cuda")
    x0 = x.clone()
    ops.gemm(A, W, bias, x, ops.EPI_GATE_RESID, gate=gate)
    assert rel(x, x0 + gate[:, None] * ref) < F32_TOL
    assert rel(x[:, -256:], (x0 + gate[:, None] * ref)[:, -256:]) < F32_TOL      # the split tail rows specifically
    wq = 1 + 0.1 * torch.randn(64, device="cuda")
    ops.gemm(A, W, bias, o, ops.EPI_QKV_NORM, norm_w0=wq, sec_cols=N)
    assert rel(o.float(), _rms(ref, wq)) < BF16_TOL
    assert rel(o[:, -256:].float(), _rms(ref, wq)[:, -256:]) < BF16_TOL


def _rms(t, w):
    t = t.reshape(*t.shape[:-1], -1, 64)
    t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    return t.reshape(*t.shape[:-2], -1)


def test_gemm_qkv_rmsnorm_and_double_normed_lr_keys(ops):
    """Fused q|k|v projection + attn.norm_q/norm_k, and the Q1 quirk: K_lr = norm_k(norm_k(k_hq))."""
    torch.manual_seed(0)
    Bq, Nq, D = 2, 256, 1536
    S = 2 * Nq + 77
    A, W, bias, ref = _mk(Bq, Nq, 1536, 3 * D)
    wq = 1 + 0.1 * torch.randn(64, device="cuda")
    wk = 1 + 0.1 * torch.randn(64, device="cuda")
    qkv = torch.zeros(Bq, S, 3 * D, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, qkv[:, :Nq], ops.EPI_QKV_NORM, norm_w0=wq, norm_w1=wk, sec_cols=D, dbl_row_offset=Nq)
    rq, rk, rv = _rms(ref[..., :D], wq), _rms(ref[..., D:2 * D], wk), ref[..., 2 * D:]
    assert rel(qkv[:, :Nq, :D].float(), rq) < BF16_TOL
    assert rel(qkv[:, :Nq, D:2 * D].float(), rk) < BF16_TOL
    assert rel(qkv[:, :Nq, 2 * D:].float(), rv) < BF16_TOL
    assert rel(qkv[:, Nq:2 * Nq, D:2 * D].float(), _rms(rk, wk)) < BF16_TOL
    assert qkv[:, 2 * Nq:].abs().max() == 0 and qkv[:, Nq:2 * Nq, :D].abs().max() == 0


@pytest.mark.parametrize("B,H,S,Sq,lr", [(1, 2, 256, 256, None), (2, 3, 333, 333, None), (1, 24, 2381, 2381, (1024, 2048)),
                                         (2, 4, 2381, 2048, (1024, 2048)), (1, 2, 100, 100, None), (1, 1, 1, 1, None),
                                         (1, 2, 8525, 8525, (4096, 8192)),
                                         # many work items per persistent CTA, even / odd KV tile counts
                                         (4, 24, 700, 700, None), (8, 24, 600, 333, (100, 300)), (3, 24, 129, 129, None)])
def test_joint_attention(ops, B, H, S, Sq, lr):
    torch.manual_seed(0)
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda").bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    o = torch.zeros(B, Sq, D, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(B, S, D, device="cuda").bfloat16() if lr else None
    ops.attention(q, k, v, o, H, seq_q=Sq, lr_res=res, lr_rows=lr or (0, 0))

    def hd(t):
        return t.float().reshape(B, S, H, 64).transpose(1, 2)

    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, S, D)
    ref = ref[:, :Sq].clone()
    if lr:
        ref[:, lr[0]:lr[1]] += res[:, lr[0]:lr[1]].float()
    assert rel(o.float(), ref) < 4e-3   # bf16 P and bf16 output roundings


LOG2E_OVER_8 = 1.4426950408889634 / 8.0


@pytest.mark.parametrize("variant", [0, 2, 3, 4])
@pytest.mark.parametrize("B,H,S,Sq,lr", [(1, 2, 256, 256, None), (2, 3, 333, 333, None), (2, 4, 2381, 2048, (1024, 2048)),
                                         (1, 1, 1, 1, None), (3, 24, 129, 129, None)])
def test_joint_attention_every_kernel_variant(ops, B, H, S, Sq, lr, variant):
    """The three pipelines (2: rotating 128-key buffers, 3: two tiles per CTA with double-buffered 96-key tiles, 4: one tile
    per CTA, two CTAs per SM) and the default choice, online softmax, against fp32 SDPA."""
    torch.manual_seed(0)
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda").bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    o = torch.zeros(B, Sq, D, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(B, S, D, device="cuda").bfloat16() if lr else None
    ops.attention(q, k, v, o, H, seq_q=Sq, lr_res=res, lr_rows=lr or (0, 0), variant=variant)
    hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, S, D)
    ref = ref[:, :Sq].clone()
    if lr:
        ref[:, lr[0]:lr[1]] += res[:, lr[0]:lr[1]].float()
    assert rel(o.float(), ref) < 4e-3


@pytest.mark.parametrize("variant", [0, 5])
@pytest.mark.parametrize("exp_offload", [0, 1, 3])
@pytest.mark.parametrize("B,H,S,Sq,lr", [(1, 2, 256, 256, None), (2, 3, 333, 333, None), (1, 24, 2381, 2381, (1024, 2048)),
                                         (2, 4, 2381, 2048, (1024, 2048)), (1, 1, 1, 1, None), (8, 24, 600, 333, (100, 300)),
                                         (1, 2, 8525, 8525, (4096, 8192)), (1, 2, 50, 50, None), (2, 2, 2350, 2350, None)])
def test_joint_attention_prescaled_bounded_scores(ops, B, H, S, Sq, lr, exp_offload, variant):
    """DIT4SR_ATTN_PRESCALED (ABI v3): q arrives multiplied by scale*log2(e) and the caller bounds every score; the
    kernel takes p = 2^(q.k) with no running max.  Reference: fp32 SDPA on the SAME (bf16-rounded, pre-scaled) q with
    scale = ln 2, i.e. softmax(2^(q'.k))."""
    torch.manual_seed(0)
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda")
    qkv[..., :D] *= LOG2E_OVER_8
    qkv = qkv.bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]

    def hd(t):
        return t.float().reshape(B, S, H, 64).transpose(1, 2)

    bound = float((hd(q)[:, :, :Sq] @ hd(k).transpose(-1, -2)).max()) if S <= 2381 else 30.0
    assert bound < 60
    o = torch.zeros(B, Sq, D, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(B, S, D, device="cuda").bfloat16() if lr else None
    ops.attention(q, k, v, o, H, seq_q=Sq, lr_res=res, lr_rows=lr or (0, 0), score_bound=max(bound, 1.0),
                  exp_offload=exp_offload, variant=variant)
    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v), scale=math.log(2.0))
    ref = ref.transpose(1, 2).reshape(B, S, D)[:, :Sq].clone()
    if lr:
        ref[:, lr[0]:lr[1]] += res[:, lr[0]:lr[1]].float()
    assert rel(o.float(), ref) < 4e-3


def test_attention_prescaled_large_and_small_scores_stay_finite(ops):
    """Scores spread over [-55, +55] log2 units (the contract allows up to 60): no overflow, and rows whose scores are
    all very negative still normalise (2^-55 is far inside bf16 / fp32 range)."""
    torch.manual_seed(0)
    B, H, S = 1, 2, 700
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda")
    qkv[..., :D] *= 0.85
    qkv[0, :100, :D] = -qkv[0, :100, D:2 * D].mean(dim=0, keepdim=True) * 6.0   # rows that dislike every key
    qkv = qkv.bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
    sc = hd(q) @ hd(k).transpose(-1, -2)
    assert float(sc.max()) < 60 and float(sc.max()) > 20
    o = torch.zeros(B, S, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, o, H, score_bound=60.0)
    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v), scale=math.log(2.0))
    ref = ref.transpose(1, 2).reshape(B, S, D)
    assert torch.isfinite(o.float()).all()
    assert rel(o.float(), ref) < 6e-3


@pytest.mark.parametrize("prescaled", [False, True])
def test_attention_at_the_c4_per_rank_shape(ops, prescaled):
    """BASELINE config C4 on 8 ranks: every rank runs 3 heads over the whole S = 33101 joint sequence (345 KV tiles of 96).
    fp32 reference by query chunks (the full score matrix would not fit comfortably)."""
    torch.manual_seed(0)
    B, H, S = 1, 3, 33101
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda")
    if prescaled:
        qkv[..., :D] *= LOG2E_OVER_8
    qkv = qkv.bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    o = torch.zeros(B, S, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, o, H, score_bound=30.0 if prescaled else None)
    hd = lambda t: t.float().reshape(B, -1, H, 64).transpose(1, 2)
    worst = 0.0
    for s0 in range(0, S, 4096):
        ref = torch.nn.functional.scaled_dot_product_attention(
            hd(q[:, s0:s0 + 4096]), hd(k), hd(v), scale=math.log(2.0) if prescaled else None)
        ref = ref.transpose(1, 2).reshape(B, -1, D)
        worst = max(worst, rel(o[:, s0:s0 + 4096].float(), ref))
    assert worst < 4e-3


@pytest.mark.parametrize("variant", [3, 4, 5])
@pytest.mark.parametrize("prescaled", [False, True])
@pytest.mark.parametrize("split_kv", [0, 2, 3, 4])
@pytest.mark.parametrize("B,H,S,Sq,lr", [(1, 24, 2381, 2381, (1024, 2048)),      # 240 pair items on 148 SMs: 92 tail items
                                         (2, 3, 4000, 3900, None),                 # 96 items: everything is tail
                                         (1, 2, 1000, 1000, None)])
def test_attention_split_kv_tail(ops, B, H, S, Sq, lr, split_kv, prescaled, variant):
    """Work items that do not fill the last wave of SMs are cut into KV slices whose partial rows (unnormalised O, row sum,
    reference max) are merged by a second small kernel: same result as the unsplit kernel, for both softmax forms, ragged
    last tiles and the LR residual (two launches on the same scratch)."""
    torch.manual_seed(0)
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda")
    if prescaled:
        qkv[..., :D] *= LOG2E_OVER_8
    else:
        qkv[..., D:2 * D] *= torch.linspace(0.3, 3.0, S, device="cuda")[None, :, None]   # maxima differ between slices
    qkv = qkv.bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    res = torch.randn(B, S, D, device="cuda").bfloat16() if lr else None
    ws = torch.empty(ops.attention_workspace_bytes(), device="cuda", dtype=torch.uint8)
    hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v), scale=math.log(2.0) if prescaled else None)
    ref = ref.transpose(1, 2).reshape(B, S, D)[:, :Sq].clone()
    if lr:
        ref[:, lr[0]:lr[1]] += res[:, lr[0]:lr[1]].float()
    for _ in range(2):
        o = torch.zeros(B, Sq, D, device="cuda", dtype=torch.bfloat16)
        ops.attention(q, k, v, o, H, seq_q=Sq, lr_res=res, lr_rows=lr or (0, 0), score_bound=30.0 if prescaled else None,
                      variant=variant, workspace=ws, split_kv=split_kv)
        err = rel(o.float(), ref)
        assert err < (4e-3 if prescaled else 6e-3), err


def test_attention_mode_is_validated(ops):
    q = torch.zeros(1, 128, 64, device="cuda", dtype=torch.bfloat16)
    o = torch.zeros_like(q)
    with pytest.raises(Exception, match="score_bound"):
        ops.attention(q, q, q, o, 1, score_bound=61.0)
    with pytest.raises(Exception, match="exp_offload"):
        ops.attention(q, q, q, o, 1, exp_offload=7)


def test_attention_peaked_softmax_is_stable(ops):
    """Large score range (|s| up to ~60): online softmax with running max must not overflow."""
    torch.manual_seed(0)
    B, H, S = 1, 2, 700
    D = H * 64
    qkv = (torch.randn(B, S, 3 * D, device="cuda") * 3).bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    o = torch.zeros(B, S, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, o, H)
    hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, S, D)
    assert torch.isfinite(o.float()).all()
    assert rel(o.float(), ref) < 6e-3


@pytest.mark.parametrize("S", [640, 1101])
def test_attention_scores_growing_along_keys(ops, S):
    """Key norms grow along the sequence, so in every KV tile some rows outgrow their reference max by more than the
    lazy-rescale threshold: exercises the fall-back of the speculative softmax (exponentials taken against the
    previous reference, redone when the vote says a row grew) on whole and ragged last tiles."""
    torch.manual_seed(0)
    B, H = 2, 3
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda")
    qkv[..., D:2 * D] *= torch.linspace(0.05, 6.0, S, device="cuda")[None, :, None]
    qkv = qkv.bfloat16()
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    o = torch.zeros(B, S, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, o, H)
    hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, S, D)
    assert torch.isfinite(o.float()).all()
    assert rel(o.float(), ref) < 6e-3


def test_attention_rows_of_ones_at_full_size(ops):
    """Size-independent property at the C2 shape (B=8, 24 heads, S=2381): with V = 1 every output element is
    sum(p)/sum(p) = 1 whatever the scores are, and exactly 1 + residual on the LR rows."""
    torch.manual_seed(0)
    B, H, N, T = 8, 24, 1024, 333
    S, D = 2 * N + T, H * 64
    qkv = torch.randn(B, S, 3 * D, device="cuda").bfloat16()
    qkv[..., 2 * D:] = 1.0
    res = torch.randn(B, 2 * N, D, device="cuda").bfloat16()
    o = torch.zeros(B, S, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], o, H, lr_res=res, lr_rows=(N, 2 * N))
    ref = torch.ones(B, S, D, device="cuda")
    ref[:, N:2 * N] += res[:, N:].float()
    assert (o.float() - ref).abs().max().item() <= 2.0 ** -5      # bf16 rounding of the output (|1 + res| < 8)


def test_head_tail_kernels(ops):
    torch.manual_seed(0)
    x = torch.randn(2, 16, 8, 12, device="cuda")
    p = torch.empty(2, 24, 64, device="cuda", dtype=torch.bfloat16)
    ops.patchify(x, p)
    ref = x.reshape(2, 16, 4, 2, 6, 2).permute(0, 2, 4, 1, 3, 5).reshape(2, 24, 64)
    assert torch.equal(p, ref.bfloat16())
    tok = torch.randn(2, 24, 64, device="cuda")
    o = torch.empty(2, 16, 8, 12, device="cuda")
    ops.unpatchify(tok, o)
    ref = tok.reshape(2, 4, 6, 2, 2, 16).permute(0, 5, 1, 3, 2, 4).reshape(2, 16, 8, 12)   # einops pattern of :391
    assert torch.equal(o, ref)
    lat = torch.randn(2, 16, 8, 12, device="cuda")
    ops.unpatchify(tok, o, latents=lat, dsigma=-0.025)
    assert rel(o, lat - 0.025 * ref) < 1e-6
    t = torch.tensor([1000.0, 741.44, 8.93], device="cuda")
    f = torch.empty(3, 256, device="cuda", dtype=torch.bfloat16)
    ops.timestep_features(t, f)
    ex = torch.exp(-math.log(10000) * torch.arange(128, device="cuda", dtype=torch.float32) / 128)
    e = t[:, None] * ex[None]
    assert (f.float() - torch.cat([torch.cos(e), torch.sin(e)], -1)).abs().max() < 8e-3


def test_errors_are_reported_not_swallowed(ops):
    from dit4sr_b200._lib import Dit4srError
    A = torch.randn(4, 20, device="cuda").bfloat16()      # K=20 not a multiple of 8
    W = torch.randn(32, 20, device="cuda").bfloat16()
    o = torch.empty(4, 32, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(Dit4srError):
        ops.gemm(A, W, None, o)
    with pytest.raises(Dit4srError):
        ops.gemm(A.cpu(), W, None, o)
