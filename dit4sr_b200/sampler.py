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


@torch.no_grad()
def denoise(model, latents, controlnet_image, encoder_hidden_states, pooled_projections, num_inference_steps=40,
            sigmas=None):
    """Runs the full sampling loop on the CUDA engine; returns the final latents (same dtype as `latents`).

    latents are carried in fp32 between steps (scheduler.step upcasts to fp32, pipeline_dit4sr.py:1567)."""
    eng = model.engine()
    sig = flow_match_sigmas(num_inference_steps) if sigmas is None else np.asarray(sigmas, dtype=np.float32)
    B = latents.shape[0]
    sp = eng.sp
    if sp is not None:   # Ulysses: every rank carries its own row strip through the loop; ONE gather at the end
        g0, g1 = sp.row_slice(latents.shape[2] // 2)
        latents_l = latents[:, :, 2 * g0:2 * g1]
        controlnet_image = controlnet_image[:, :, 2 * g0:2 * g1].contiguous()
    else:
        latents_l = latents
    x = latents_l.float().contiguous()
    nxt = torch.empty_like(x)
    t = torch.empty(B, device=latents.device, dtype=torch.float32)
    for i in range(len(sig) - 1):
        t.fill_(float(sig[i]) * 1000.0)
        eng.step(x, controlnet_image, encoder_hidden_states, pooled_projections, t, out=nxt,
                 euler=(x, float(sig[i + 1]) - float(sig[i])), sharded_io=sp is not None)
        x, nxt = nxt, x
    if sp is not None:
        x = sp.gather_rows(x, dim=2)
    return x.to(latents.dtype)
