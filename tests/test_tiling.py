"""Latent tiling (row N3): the batched tiling of dit4sr_b200.tiling vs the literal restatement of the reference loops."""
import pytest
import torch

from dit4sr_b200.tiling import gaussian_weights, tile_offsets, tiled_velocity
from oracle import dit4sr_oracle as O


def test_offsets_cover_and_end_flush():
    for size, tile, ov in [(64, 24, 8), (40, 24, 8), (24, 24, 8), (100, 32, 4), (33, 32, 16)]:
        offs = tile_offsets(size, tile, ov)
        assert offs[0] == 0 and offs[-1] == size - tile
        covered = set()
        for o in offs:
            covered.update(range(o, o + tile))
        assert covered == set(range(size))


@pytest.mark.parametrize("h,w,tile,ov", [(24, 24, 16, 4), (16, 28, 16, 8), (16, 16, 16, 4)])
def test_batched_tiling_equals_reference_loops(h, w, tile, ov):
    cfg = O.tiny_config()
    sd = O.init_state_dict(cfg)
    g = torch.Generator().manual_seed(3)
    lat, ctl = torch.randn(1, 16, h, w, generator=g), torch.randn(1, 16, h, w, generator=g)
    _, _, enc, pooled, t = O.make_inputs(cfg, 1, 8, 8, 5)

    def model_fn(x, c):
        n = x.shape[0]
        return O.model_forward(sd, cfg, x, c, enc.expand(n, -1, -1), pooled.expand(n, -1), t.expand(n), as_written=False)

    ref = O.tiled_velocity_as_written(model_fn, lat, ctl, tile, ov)
    mine = tiled_velocity(model_fn, lat, ctl, tile, ov)
    mine2 = tiled_velocity(model_fn, lat, ctl, tile, ov, tile_batch=2)
    assert (mine - ref).abs().max() <= 2e-5
    assert (mine2 - ref).abs().max() <= 2e-5


def test_gaussian_window_matches_reference_formula():
    w = gaussian_weights(8, 8)
    assert w.shape == (8, 8) and (w > 0).all()
    assert torch.argmax(w[4]).item() in (3, 4)


@pytest.mark.gpu
def test_tiled_velocity_on_gpu_matches_oracle():
    import dit4sr_b200
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
    m.load_state_dict(sd)
    m = m.to("cuda", torch.bfloat16).requires_grad_(False).eval()
    g = torch.Generator().manual_seed(3)
    lat, ctl = torch.randn(1, 16, 48, 48, generator=g), torch.randn(1, 16, 48, 48, generator=g)
    _, _, enc, pooled, t = O.make_inputs(cfg, 1, 8, 8, 20, timestep=400.0)
    bf = torch.bfloat16

    def gpu_fn(x, c):
        n = x.shape[0]
        with torch.no_grad():
            return m(x.to(bf), c.to(bf), enc.cuda().to(bf).expand(n, -1, -1).contiguous(),
                     pooled.cuda().to(bf).expand(n, -1).contiguous(), t.cuda().expand(n), return_dict=False)[0]

    def cpu_fn(x, c):
        n = x.shape[0]
        return O.model_forward(sd, cfg, x, c, enc.expand(n, -1, -1), pooled.expand(n, -1), t.expand(n), as_written=False)

    out = tiled_velocity(gpu_fn, lat.cuda(), ctl.cuda(), 32, 8).cpu()
    ref = O.tiled_velocity_as_written(cpu_fn, lat, ctl, 32, 8)
    assert ((out - ref).norm() / ref.norm()).item() <= 1e-2


@pytest.mark.gpu
def test_tiled_sampling_loop_of_a_non_square_latent():
    """denoise(..., latent_tiled_size, latent_tiled_overlap) on a 40x56 latent (a 320x448 image): three Euler steps, every
    step's velocity blended from overlapping 32x32 tiles run as one batch, against the oracle's as-written tile loop."""
    import dit4sr_b200
    from dit4sr_b200.sampler import denoise
    cfg = O.sd35_medium_config(num_layers=2, dual_attention_layers=(0,), pos_embed_max_size=96)
    sd = O.init_state_dict(cfg, seed=0)
    m = dit4sr_b200.SD3Transformer2DModel(**vars(cfg))
    m.load_state_dict(sd)
    m = m.to("cuda", torch.bfloat16).requires_grad_(False).eval()
    g = torch.Generator().manual_seed(5)
    lat, ctl = torch.randn(1, 16, 40, 56, generator=g), torch.randn(1, 16, 40, 56, generator=g)
    _, _, enc, pooled, _ = O.make_inputs(cfg, 1, 8, 8, 20)
    bf = torch.bfloat16
    out = denoise(m, lat.cuda(), ctl.cuda(), enc.cuda().to(bf), pooled.cuda().to(bf), num_inference_steps=3,
                  latent_tiled_size=32, latent_tiled_overlap=8).float().cpu()
    sig = O.flow_match_sigmas(3)
    x = lat.clone()
    for i in range(3):
        t = torch.full((1,), float(sig[i]) * 1000.0)

        def cpu_fn(xt, ct):
            n = xt.shape[0]
            return O.model_forward(sd, cfg, xt, ct, enc.expand(n, -1, -1), pooled.expand(n, -1), t.expand(n), as_written=False)

        x = x + (float(sig[i + 1]) - float(sig[i])) * O.tiled_velocity_as_written(cpu_fn, x, ctl, 32, 8)
    assert out.shape == x.shape
    assert ((out - x).norm() / x.norm()).item() <= 1e-2


def test_tile_plan_is_a_partition_of_unity_and_orders_tiles_like_the_reference():
    """Size-independent property: blending the tiles cut from a tensor gives the tensor back (the gaussian weights are
    normalised by their sum), for square, non-square and barely-overlapping layouts; tile order = the reference's loop
    order (x outer, y inner, pipeline_dit4sr.py:1488-1492) with batch index = tile * B + image."""
    from dit4sr_b200.tiling import TilePlan
    torch.manual_seed(0)
    for (h, w, tile, ov) in [(128, 128, 64, 24), (40, 56, 32, 8), (64, 64, 64, 24), (65, 129, 64, 1), (24, 24, 64, 8)]:
        plan = TilePlan(h, w, tile, ov)
        x = torch.randn(3, 4, h, w)
        tiles = plan.extract(x)
        T = min(tile, h, w)
        assert tiles.shape == (3 * len(plan), 4, T, T)
        assert torch.allclose(plan.blend(tiles), x, atol=1e-5)
        assert plan.weight_sum.min() > 0
        y0, x0 = plan.coords[1] if len(plan) > 1 else plan.coords[0]
        k = 1 if len(plan) > 1 else 0
        assert torch.equal(tiles[3 * k + 2], x[2, :, y0:y0 + T, x0:x0 + T])
        xs = sorted({c[1] for c in plan.coords})
        assert [c[1] for c in plan.coords] == [xx for xx in xs for _ in range(len(plan) // len(xs))]
