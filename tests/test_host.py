"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, drop-in module surface."""
import ctypes
import os
import re

import pytest
import torch

import dit4sr_b200
from dit4sr_b200 import _lib, build
from oracle import dit4sr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dit4sr_b200.h")).read()
    declared = set(re.findall(r"\b(dit4sr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dit4sr_abi_version() == 5


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    p = _lib.GemmParams()
    rc = lib.dit4sr_gemm_bf16(ctypes.byref(p), None)
    assert rc == -1 and b"null pointer" in lib.dit4sr_last_error()
    rc = lib.dit4sr_patchify(None, 0, None, 1, 16, 8, 8, None)
    assert rc == -1


def _check_schedule(shapes, sms):
    """Every output tile of every problem is covered exactly once, in problem order, and only the final partial wave is
    cut into column slices."""
    from dit4sr_b200 import ops
    items = ops.gemm_schedule(shapes, sm_count=sms)
    cover = {}
    for pi, b, mt, c0, bn in items:
        B, R, N, K = shapes[pi]
        assert 0 <= b < B and 0 <= mt < -(-R // 128) and 0 <= c0 < N and bn in (64, 128, 256)
        cover.setdefault((pi, b, mt), []).append((c0, bn))
    n_tiles = 0
    for pi, (B, R, N, K) in enumerate(shapes):
        for b in range(B):
            for mt in range(-(-R // 128)):
                segs = sorted(cover.pop((pi, b, mt)))
                pos = 0
                for c0, bn in segs:                  # contiguous, non-overlapping, covering [0, N)
                    assert c0 == pos
                    pos += bn
                assert pos >= N and pos - N < 256
                n_tiles += -(-N // 256)
    assert not cover
    assert [it[0] for it in items] == sorted(it[0] for it in items)      # problems scheduled in order
    sliced = [i for i, it in enumerate(items) if it[4] != 256]
    if sliced:
        rem = n_tiles % sms
        split = 256 // items[sliced[0]][4]
        assert rem > 0 and n_tiles > sms and split * rem <= sms
        assert sliced == list(range(n_tiles - rem, n_tiles - rem + split * rem))   # exactly the remainder tiles
        assert all(N % 256 == 0 for _, _, N, _ in shapes)
    else:
        assert len(items) == n_tiles
    return items


def test_gemm_schedule_of_the_model_launches_without_a_gpu():
    """The tile schedule (host arithmetic shared with the launch path) for the C2 launches: q|k|v group, out-projection
    group (912 = 6 x 148 + 24 tiles -> the 24 remainder tiles become 96 slices of 64 columns), FF1/FF2 groups."""
    B, N, T, D, F = 8, 1024, 333, 1536, 6144
    _check_schedule([(B, N, 3 * D, D), (B, N, 2 * D, D), (B, T, 3 * D, D)], 148)
    items = _check_schedule([(B, N, D, D), (B, N, D, D), (B, T, D, D)], 148)
    assert sum(1 for it in items if it[4] == 64) == 96 and len(items) == 888 + 96
    _check_schedule([(B, 2 * N, F, D), (B, T, F, D)], 148)
    _check_schedule([(B, 2 * N, D, F), (B, T, D, F)], 148)
    _check_schedule([(1, 7, 96, 64)], 148)             # a single ragged tile
    _check_schedule([(3, 333, 1504, 128)], 4)          # N not a multiple of 256: never sliced


def test_gemm_schedule_random_shapes():
    import random
    rng = random.Random(0)
    for _ in range(200):
        shapes = [(rng.randint(1, 4), rng.randint(1, 700), 32 * rng.randint(1, 40) if rng.random() < 0.5
                   else 256 * rng.randint(1, 8), 8 * rng.randint(1, 64)) for _ in range(rng.randint(1, 6))]
        _check_schedule(shapes, rng.choice([1, 4, 37, 148, 160]))


def test_gemm_schedule_rejects_bad_groups(lib):
    arr = (_lib.GemmParams * 7)()
    assert lib.dit4sr_gemm_schedule(arr, 7, 148, None, 0) == -1          # more than DIT4SR_GEMM_MAX_GROUP problems
    arr[0].batches, arr[0].rows, arr[0].N, arr[0].K = 1, 128, 48, 64     # N % 32 != 0
    assert lib.dit4sr_gemm_schedule(arr, 1, 148, None, 0) < 0


def test_state_dict_layout_matches_reference_manifest():
    for cfg in (O.tiny_config(), O.tiny_config(qk_norm=None, dual_attention_layers=())):
        m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
        man = O.state_dict_manifest(cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == list(man.keys())
        assert all(tuple(sd[k].shape) == man[k] for k in sd)


def test_control_modules_zero_initialised_like_reference():
    m = dit4sr_b200.SD3Transformer2DModel(**vars(O.tiny_config()))
    a = m.transformer_blocks[0].attn
    assert a.to_q_control.weight.abs().max() == 0 and a.to_v_control.bias.abs().max() == 0
    assert torch.equal(a.to_out_control.weight, a.to_out[0].weight)
    assert m.transformer_blocks[0].ff.control_conv.weight.abs().max() == 0


def test_no_cpu_fallback():
    cfg = O.tiny_config()
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg)).requires_grad_(False)
    hs, lr, enc, pooled, t = O.make_inputs(cfg, 1, 8, 8, 5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(hs, lr, enc, pooled, t)


def test_save_and_load_roundtrip(tmp_path):
    cfg = O.tiny_config()
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
    m.load_state_dict(O.init_state_dict(cfg))
    m.save_pretrained(str(tmp_path / "transformer"))
    m2 = dit4sr_b200.SD3Transformer2DModel.from_pretrained_local(str(tmp_path), subfolder="transformer")
    assert m2.config.dual_attention_layers == (0,) and m2.config.qk_norm == "rms_norm"
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_sampler_sigmas_match_oracle_and_logged_timesteps():
    import json
    import numpy as np
    from dit4sr_b200.sampler import flow_match_sigmas
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_meta.json")))
    s = flow_match_sigmas(40)
    assert np.array_equal(s, O.flow_match_sigmas(40))
    assert np.all(np.abs(1000 * s[:-1] - np.array(meta["logged_timesteps_40"])) <= 0.005 + 1e-6)


def test_bench_attention_roofline_entry():
    """bench.py's `roofline` entry (joint attention = the largest share of the step): algorithmic FLOPs / measured time."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    peaks = dict(tf_burst=1658.4, tf_sustained=1390.8, hbm=6540.0, source="test")
    gemm_f, attn_f = bench.flops_per_image_step(1024, 333)
    r = bench.attention_roofline([0.5] * 37 + [0.4] * 37, attn_f * 8, 37, peaks)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["launches_timed"] == 74
    assert abs(r["achieved"] - attn_f * 8 / 37 / 0.45e-3 / 1e12) < 0.1
    assert abs(r["frac"] - r["achieved"] / 1390.8) < 1e-3 and r["traffic"] is None or r["traffic"] > 1e8
    assert bench.attention_roofline([], attn_f, 37, peaks) is None
    assert abs((gemm_f + attn_f) / 1e12 - 4.724) < 0.01          # folded-chain total at 512px


def test_bench_hbm_roofline_entries():
    """bench.py's `roofline_hbm`: algorithmic bytes of the LayerNorm+modulate launches and of the stencil of one C2 step
    (3 N F 2 bytes per image and block for the stencil; LayerNorm reads the fp32 stream, writes bf16) / time."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    B, N, T, D, F = 8, 1024, 333, 1536, 6144
    bd = {"ln_mod": {"launches": 48, "total_ms": 2.0}, "dwconv": {"launches": 24, "total_ms": 1.5}}
    r = bench.hbm_rooflines(bd, B, N, T, 24, 13, dict(hbm=6541.1))
    assert r["dwconv3x3"]["algorithmic_bytes_per_step"] == 24 * 3 * N * B * F * 2
    img, txt = 2 * N * B * D, T * B * D
    ln = 13 * img * 8 + 11 * img * 6 + 24 * img * 6 + (24 + 23) * txt * 6      # dual norm1, single norm1, norm2, text rows
    assert r["ln_modulate"]["algorithmic_bytes_per_step"] == ln
    assert abs(r["ln_modulate"]["achieved"] - ln / 2.0e-3 / 1e9) < 0.1 and r["ln_modulate"]["bound"] == "hbm"
    assert abs(r["dwconv3x3"]["frac"] - r["dwconv3x3"]["achieved"] / 6541.1) < 1e-3
    assert bench.hbm_rooflines(None, B, N, T, 24, 13, dict(hbm=1.0)) is None


def test_bench_product_arm_config_matches_survey_and_avoids_the_oracle():
    """bench.py builds the product-arm model from its own dict (nothing under oracle/ outside the CPU legs); that dict,
    completed with the drop-in constructor's defaults, is the oracle's SD3.5-medium config."""
    import inspect
    import sys
    sys.path.insert(0, ROOT)
    import bench
    d = {k: v.default for k, v in inspect.signature(dit4sr_b200.SD3Transformer2DModel.__init__).parameters.items()
         if k != "self"}
    d.update(bench.SD35_MEDIUM_CONFIG)
    assert d == vars(O.sd35_medium_config())
    src = inspect.getsource(bench.run_ours) + inspect.getsource(bench.build_model_gpu)
    assert "oracle" not in src


def test_every_entry_point_rejects_null_arguments_with_a_message(lib):
    """Error behaviour of the C ABI: no crash, DIT4SR_E_BADARG (-1) and a message naming the entry point, without a GPU."""
    for name, (_, args) in _lib.SYMBOLS.items():
        if name in ("dit4sr_abi_version", "dit4sr_last_error", "dit4sr_check_device",
                    "dit4sr_peer_free", "dit4sr_peer_close",       # free(NULL) / close(NULL) are no-ops by design
                    "dit4sr_attn_workspace_bytes",                 # a size query without arguments
                    "dit4sr_groupnorm_workspace_bytes"):           # a size query (0 for non-positive sizes)
            continue
        vals = [0 if a in (ctypes.c_int, ctypes.c_int64, ctypes.c_size_t) else (0.0 if a is ctypes.c_float else None) for a in args]
        rc = getattr(lib, name)(*vals)
        assert rc == -1, (name, rc)
        assert len(lib.dit4sr_last_error()) > 0, name


def test_attention_work_decomposition_without_a_gpu():
    """dit4sr_attn_plan (host only): kernel choice and the split-KV tail for the shapes of BASELINE's configs on 148 SMs."""
    from dit4sr_b200 import ops
    # C2 joint attention: 3648 query tiles on 296 slots = 12.32 waves: v4; the 96 tail tiles are NOT sliced: they run one
    # per SM at single-CTA speed, and two extra launches cost more than a third of a 22 us item (measured: 0.298 vs 0.314 ms)
    v, split, whole, tail, tiles, tps = ops.attention_plan(8, 24, 2381, 2381)
    assert (v, split, whole, tail) == (4, 1, 19 * 24 * 8, 0) and tiles == 25 and tps == 25
    assert ops.attention_plan(8, 24, 2381, 2381, workspace=False)[:4] == (4, 1, 19 * 24 * 8, 0)
    assert ops.attention_plan(8, 24, 2381, 2381, split_kv=3)[:4] == (4, 3, 12 * 296, 96)          # forced
    # pre-scaled scores: same decomposition on the split-row form (5); an explicit 5 without pre-scaled scores runs as 4
    assert ops.attention_plan(8, 24, 2381, 2381, prescaled=True)[:4] == (5, 1, 19 * 24 * 8, 0)
    assert ops.attention_plan(8, 24, 2381, 2381, variant=5)[0] == 4
    # C4 on 8 ranks: 3 heads over S = 33101: 390 pair items on 148 SMs = 2.64 waves -> v3, 94 tail items in 3 KV slices
    v, split, whole, tail, tiles, tps = ops.attention_plan(1, 3, 33101, 33101)
    assert (v, split, whole, tail) == (3, 3, 296, 94) and tiles == 345 and tps == 115
    # the same without scratch: no split, a pair-per-CTA kernel (long sequence: K/V sharing beats prologue hiding)
    v, split, whole, tail, _, _ = ops.attention_plan(1, 3, 33101, 33101, workspace=False)
    assert v in (2, 3) and (split, whole, tail) == (1, 390, 0)
    # invariants over a sweep of shapes (and forced splits)
    for B, H, Sq, S in [(1, 1, 1, 1), (2, 3, 3900, 4000), (8, 24, 2048, 2048), (2, 24, 8525, 8525), (1, 24, 32768, 33101),
                        (4, 24, 8192, 8525), (1, 2, 100, 100), (3, 24, 129, 129)]:
        for force in (0, 2, 3, 4):
            for variant in (0, 3, 4, 5):
                v, split, whole, tail, tiles, tps = ops.attention_plan(B, H, Sq, S, split_kv=force, variant=variant,
                                                                       prescaled=variant == 5)
                assert variant == 0 or v == variant
                rows = 128 if v >= 4 else 256
                items = -(-Sq // rows) * H * B
                assert whole + tail == items and 1 <= split <= 4
                assert (split == 1) == (tail == 0)
                assert tiles == -(-S // (128 if v == 2 else 96))
                assert tps * split >= tiles and tps * (split - 1) < tiles      # every slice has at least one KV tile
                slots = 296 if v >= 4 else 148
                if split > 1:
                    assert whole % slots == 0 and tail < slots


def test_sampler_forms_the_loop_table_once_and_hands_every_step_its_slice():
    """Host logic of sampler.denoise around Engine.prepare_loop (no GPU: a recording stand-in for the engine): one table per
    loop built from the doubled prompt batch and the schedule's timesteps, step i called with loop=(cache, i); with
    hoist_loop_invariants=False (or when prepare_loop declines) every step recomputes its head (loop=None)."""
    from dit4sr_b200 import sampler

    class FakeEngine:
        sp = None

        def __init__(self, decline=False):
            self.decline, self.tables, self.steps = decline, [], []

        def prepare_loop(self, enc, pooled, timesteps):
            self.tables.append((tuple(enc.shape), tuple(pooled.shape), [round(float(t), 3) for t in timesteps]))
            return None if self.decline else {"id": len(self.tables)}

        def step(self, x, control, enc, pooled, t, out=None, euler=None, sharded_io=False, guidance=None, loop=None):
            self.steps.append((loop, float(t[0]), guidance, x.shape[0]))
            out.copy_(euler[0])

    class FakeModel:
        def __init__(self, eng):
            self._eng = eng

        def engine(self):
            return self._eng

    lat, ctl = torch.randn(2, 16, 8, 8), torch.randn(2, 16, 8, 8)
    enc, pooled = torch.randn(2, 5, 32), torch.randn(2, 16)
    sig = sampler.flow_match_sigmas(4)
    eng = FakeEngine()
    sampler.denoise(FakeModel(eng), lat, ctl, enc, pooled, num_inference_steps=4)
    assert len(eng.tables) == 1 and eng.tables[0][0] == (2, 5, 32)
    assert eng.tables[0][2] == [round(float(s) * 1000.0, 3) for s in sig[:-1]]
    assert [s[0][1] for s in eng.steps] == [0, 1, 2, 3] and all(s[0][0] == {"id": 1} for s in eng.steps)
    assert [round(s[1], 3) for s in eng.steps] == eng.tables[0][2]
    eng = FakeEngine()
    sampler.denoise(FakeModel(eng), lat, ctl, enc, pooled, num_inference_steps=4, guidance_scale=8.0,
                    negative_prompt_embeds=torch.zeros_like(enc), negative_pooled_prompt_embeds=torch.zeros_like(pooled))
    assert eng.tables[0][0] == (4, 5, 32) and eng.tables[0][1] == (4, 16)       # [negative | positive] prompt batch
    assert all(s[2] == 8.0 and s[3] == 4 for s in eng.steps)
    for kw, eng in ((dict(hoist_loop_invariants=False), FakeEngine()), (dict(), FakeEngine(decline=True))):
        sampler.denoise(FakeModel(eng), lat, ctl, enc, pooled, num_inference_steps=3, **kw)
        assert all(s[0] is None for s in eng.steps) and len(eng.steps) == 3
