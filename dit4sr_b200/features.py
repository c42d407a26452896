"""Feature outputs for the text-spotting module (`ts_module`), the consumer of the denoiser's per-block hidden states.

The reference pipeline turns every returned `extract_feat` slice [B, N*1024, 1536] into an image-shaped fp32 tensor with
    rearrange(feat, 'b (N H W) (pH pW d) -> b (N d) (H pH) (W pW)', N=num_concat_feat, H=32, W=32, pH=2, pW=2).to(float32)
(pipelines/pipeline_dit4sr.py:1214-1245).  `unpatchify_features` is that step: it returns the tensors the CUDA engine already
wrote while the block's residual stream was at hand (`extract_feat_nchw_f32`, one kernel inside the captured graph), and runs
the same kernel on a bare `extract_feat` otherwise."""
import torch

from . import ops


def unpatchify_features(etc_out, feat_extract: str = "hq_feat", patch_size: int = 2, latent_size: int = 64):
    """etc_out: the list of dicts `SD3Transformer2DModel.forward(..., cfg=cfg, return_dict=False)[1]` returns.
    Returns the list of fp32 [B, N*d, 64, 64] tensors (N = 2 for 'hqlq_feat', else 1; d = hidden / 4)."""
    if patch_size != 2:
        raise ValueError("dit4sr_b200: patch_size must be 2")
    n = 2 if feat_extract == "hqlq_feat" else 1
    g = latent_size // patch_size
    outs = []
    for d in etc_out:
        if "extract_feat_nchw_f32" in d:
            outs.append(d["extract_feat_nchw_f32"])
            continue
        f = d["extract_feat"]
        B, rows, D = f.shape
        if rows != n * g * g:
            raise ValueError(f"extract_feat has {rows} rows, expected {n}*{g}*{g}")
        nchw = torch.empty(B, n * D // 4, 2 * g, 2 * g, device=f.device, dtype=torch.float32)
        ops.feature_extract(f.float().contiguous(), 0, n, g, g, nchw=nchw)
        outs.append(nchw)
    return outs
