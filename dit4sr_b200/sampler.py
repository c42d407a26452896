"""The denoise loop around the model (pipelines/pipeline_dit4sr.py:1004-1034,1078-1132,1561-1567): flow-match Euler
schedule (FlowMatchEulerDiscreteScheduler, shift 3.0), guidance scale 1, start point 'noise', untiled branch.
The Euler update `x += (sigma_next - sigma) * v` is fused into the model's unpatchify kernel."""
import numpy as np
import torch


def flow_match_sigmas(num_steps: int, shift: float = 3.0, num_train_timesteps: int = 1000) -> np.ndarray:
    """scheduler.set_timesteps(num_steps): returns num_steps+1 sigmas (last = 0); timesteps = 1000 * sigmas[:-1]."""
    t = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
    s = t / num_train_timesteps
    s = shift * s / (1 + (shift - 1) * s)
    ts = np.linspace(float(s[0]) * num_train_timesteps, float(s[-1]) * num_train_timesteps, num_steps)
    sig = ts / num_train_timesteps
    sig = shift * sig / (1 + (shift - 1) * sig)
    return np.concatenate([sig, np.zeros(1)]).astype(np.float32)


def lre_start_point(latents, controlnet_image, sigma0: float):
    """start_point == 'lr' (the LRE strategy), pipelines/pipeline_dit4sr.py:1025-1034: the noise is blended with the
    control latent, (1 - sigma_0) * control[:1] + sigma_0 * noise - with the FIRST control latent for every image of the
    batch, as the reference writes it."""
    return (1.0 - sigma0) * controlnet_image[:1].to(latents.dtype) + sigma0 * latents


@torch.no_grad()
def denoise(model, latents, controlnet_image, encoder_hidden_states, pooled_projections, num_inference_steps=40,
            sigmas=None, guidance_scale: float = 1.0, negative_prompt_embeds=None, negative_pooled_prompt_embeds=None,
            start_point: str = "noise", latent_tiled_size: int = 0, latent_tiled_overlap: int = 0, tile_batch: int = 0,
            hoist_loop_invariants: bool = True):
    """Runs the full sampling loop on the CUDA engine; returns the final latents (same dtype as `latents`).

    latents are carried in fp32 between steps (scheduler.step upcasts to fp32, pipeline_dit4sr.py:1567).
    guidance_scale > 1 (classifier-free guidance, pipeline_dit4sr.py:1085,1106-1108,1561-1563): every step runs the model
    on the doubled batch [uncond | text] with [negative | positive] prompt embeddings and the combine
    v_uncond + g (v_text - v_uncond) is fused with the Euler update into the model's last kernel.
    start_point: 'noise' or 'lr' (LRE strategy, :1025-1034).
    latent_tiled_size / latent_tiled_overlap (:794-795, :1042): latents with more than tile^2 pixels (:1088) - the only way
    non-square images reach the model, whose token grid is square (attention.py:169-172) - are denoised from overlapping
    square tiles blended with the gaussian window (`tiling.tiled_velocity`, :1449-1558); all tiles of a step run as ONE
    batch (tile_batch > 0 caps the tiles per model call)."""
    h_lat, w_lat = latents.shape[-2:]
    if latent_tiled_size > 0 and h_lat * w_lat > latent_tiled_size * latent_tiled_size:
        return _denoise_tiled(model, latents, controlnet_image, encoder_hidden_states, pooled_projections,
                              num_inference_steps, sigmas, guidance_scale, start_point, latent_tiled_size,
                              latent_tiled_overlap, tile_batch)
    eng = model.engine()
    sig = flow_match_sigmas(num_inference_steps) if sigmas is None else np.asarray(sigmas, dtype=np.float32)
    B = latents.shape[0]
    if start_point == "lr":
        latents = lre_start_point(latents, controlnet_image, float(sig[0]))
    elif start_point != "noise":
        raise ValueError("start_point must be 'noise' or 'lr'")
    cfg_on = guidance_scale > 1.0                                  # do_classifier_free_guidance
    if cfg_on:
        if negative_prompt_embeds is None or negative_pooled_prompt_embeds is None:
            raise ValueError("guidance_scale > 1 needs negative_prompt_embeds and negative_pooled_prompt_embeds")
        encoder_hidden_states = torch.cat([negative_prompt_embeds, encoder_hidden_states], dim=0)
        pooled_projections = torch.cat([negative_pooled_prompt_embeds, pooled_projections], dim=0)
        if controlnet_image.shape[0] == B:                         # prepare_image doubles it (:714-715)
            controlnet_image = torch.cat([controlnet_image] * 2, dim=0)
    if controlnet_image.dtype not in (torch.float32, torch.bfloat16):
        controlnet_image = controlnet_image.float()                # fp16 modules: the engine takes fp32 / bf16 images
    sp = eng.sp
    if sp is not None:   # Ulysses: every rank carries its own row strip through the loop; ONE gather at the end
        g0, g1 = sp.row_slice(latents.shape[2] // 2)
        latents_l = latents[:, :, 2 * g0:2 * g1]
        controlnet_image = controlnet_image[:, :, 2 * g0:2 * g1].contiguous()
    else:
        latents_l = latents
    x = latents_l.float().contiguous()
    nxt = torch.empty_like(x)
    x2 = torch.empty((2 * B,) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype) if cfg_on else None
    t = torch.empty(2 * B if cfg_on else B, device=latents.device, dtype=torch.float32)
    # loop-invariant work hoisted out of the loop: the adaLN vectors of all steps (one GEMM over steps x batch rows) and the
    # embedded text tokens (Engine.prepare_loop); None = recompute per step
    cache = eng.prepare_loop(encoder_hidden_states, pooled_projections, [float(s) * 1000.0 for s in sig[:-1]]) \
        if hoist_loop_invariants else None
    for i in range(len(sig) - 1):
        t.fill_(float(sig[i]) * 1000.0)
        if cfg_on:
            x2[:B].copy_(x)
            x2[B:].copy_(x)
        eng.step(x2 if cfg_on else x, controlnet_image, encoder_hidden_states, pooled_projections, t, out=nxt,
                 euler=(x, float(sig[i + 1]) - float(sig[i])), sharded_io=sp is not None,
                 guidance=float(guidance_scale) if cfg_on else None, loop=None if cache is None else (cache, i))
        x, nxt = nxt, x
    if sp is not None:
        x = sp.gather_rows(x, dim=2)
    return x.to(latents.dtype)


@torch.no_grad()
def super_resolve(model, vae, lr_image, encoder_hidden_states, pooled_projections, num_inference_steps=40, generator=None,
                  latents=None, guidance_scale: float = 1.0, negative_prompt_embeds=None,
                  negative_pooled_prompt_embeds=None, start_point: str = "noise", sample_posterior: bool = True,
                  output_type: str = "pt", latent_tiled_size: int = 0, latent_tiled_overlap: int = 0, tile_batch: int = 0):
    """The image path of the reference's pipeline call around the loop (pipelines/pipeline_dit4sr.py:1001-1002, 1004-1034,
    1561-1567, 1602-1605), every stage on the CUDA library:
        control = (vae.encode(lr_image).latent_dist.sample() - shift_factor) * scaling_factor
        latents = 40-step flow-match Euler loop of the model (`denoise`)
        image   = vae.decode(latents / scaling_factor + shift_factor); postprocess: (image / 2 + 0.5).clamp(0, 1)
    lr_image: [B, 3, H, W] in [-1, 1] (what prepare_image / VaeImageProcessor.preprocess hand over), H and W multiples
    of 16; non-square or large images need latent_tiled_size (the shipped configs use 64 with overlap 24).  latents: optional start noise [B, 16, H/8, W/8] (drawn from `generator` otherwise).  output_type "pt" returns
    the image tensor in [0, 1], "latent" the final latents (pipeline_dit4sr.py:1598-1599)."""
    cfg = vae.config
    post = vae.encode(lr_image).latent_dist
    control = post.sample(generator=generator) if sample_posterior else post.mode()
    control = (control.float() - cfg.shift_factor) * cfg.scaling_factor
    dt = model.dtype
    if latents is None:
        latents = torch.randn(control.shape, generator=generator, device=control.device, dtype=torch.float32)
    final = denoise(model, latents.to(dt), control.to(dt), encoder_hidden_states, pooled_projections,
                    num_inference_steps=num_inference_steps, guidance_scale=guidance_scale,
                    negative_prompt_embeds=negative_prompt_embeds, negative_pooled_prompt_embeds=negative_pooled_prompt_embeds,
                    start_point=start_point, latent_tiled_size=latent_tiled_size, latent_tiled_overlap=latent_tiled_overlap,
                    tile_batch=tile_batch)
    if output_type == "latent":
        return final
    image = vae.decode(final.float() / cfg.scaling_factor + cfg.shift_factor, return_dict=False)[0]
    return (image / 2 + 0.5).clamp(0, 1)


def _denoise_tiled(model, latents, controlnet_image, encoder_hidden_states, pooled_projections, num_inference_steps, sigmas,
                   guidance_scale, start_point, tile, overlap, tile_batch):
    """The tiled branch of the loop (pipeline_dit4sr.py:1449-1558).  Everything that does not change from step to step is
    formed once: tile placement, window and weight sum (`TilePlan`), the control tiles, the prompt embeddings repeated per
    tile, and the engine's loop table (adaLN vectors of every step, embedded text tokens) per model-call size."""
    from . import ops
    from .tiling import TilePlan
    if guidance_scale > 1.0:
        raise NotImplementedError("tiled sampling runs at guidance_scale 1 (the reference's shipped tiled configs)")
    sig = flow_match_sigmas(num_inference_steps) if sigmas is None else np.asarray(sigmas, dtype=np.float32)
    B = latents.shape[0]
    if start_point == "lr":
        latents = lre_start_point(latents, controlnet_image, float(sig[0]))
    elif start_point != "noise":
        raise ValueError("start_point must be 'noise' or 'lr'")
    dt = model.dtype
    eng = model.engine()
    h, w = latents.shape[-2:]
    plan = TilePlan(h, w, tile, overlap, device=latents.device)
    n_tiles = len(plan)
    per_call = min(tile_batch, n_tiles) if tile_batch > 0 else n_tiles          # tiles per model call
    if dt not in (torch.float32, torch.bfloat16):
        dt = torch.float32                                                        # fp16 modules: the engine takes fp32 / bf16 images
    ctl_tiles = plan.extract(controlnet_image.to(dt)).contiguous()                # batch index = tile * B + image
    calls = [(i * B, min(i + per_call, n_tiles) * B) for i in range(0, n_tiles, per_call)]
    enc_rep = encoder_hidden_states.to(dt).repeat(per_call, 1, 1)
    pooled_rep = pooled_projections.to(dt).repeat(per_call, 1)
    timesteps = [float(s) * 1000.0 for s in sig[:-1]]
    caches = {}                                                                   # rows of a call -> loop table
    for a, b in calls:
        if b - a not in caches:
            caches[b - a] = eng.prepare_loop(enc_rep[:b - a], pooled_rep[:b - a], timesteps)
    x = latents.float().contiguous()
    nxt = torch.empty_like(x)
    for i in range(len(sig) - 1):
        lat_tiles = plan.extract(x.to(dt)).contiguous()
        outs = []
        for a, b in calls:
            t = torch.full((b - a,), timesteps[i], device=x.device, dtype=torch.float32)
            cache = caches[b - a]
            outs.append(eng.step(lat_tiles[a:b], ctl_tiles[a:b], enc_rep[:b - a], pooled_rep[:b - a], t,
                                 loop=None if cache is None else (cache, i)))
        v = plan.blend(torch.cat(outs, dim=0).float()).contiguous()
        ops.euler_update(x, v, nxt, float(sig[i + 1]) - float(sig[i]))
        x, nxt = nxt, x
    return x.to(latents.dtype)
