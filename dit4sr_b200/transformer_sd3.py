"""Drop-in for `model_dit4sr/transformer_sd3.py::SD3Transformer2DModel` of jinlovespho/DiT4SR.

Same constructor kwargs/defaults (transformer_sd3.py:64-81), same `forward` signature (:267-279), same `state_dict`
layout and the same attributes the reference's callers touch (`.config`, `.dtype`, `.transformer_blocks`,
`from_config`, `from_pretrained_local`, `save_pretrained`) -- so `pipelines/pipeline_dit4sr.py:1122-1131`
can use it unchanged.  The forward pass runs entirely in the sm_100a CUDA library (`engine.py` -> C-ABI); there is no
PyTorch/CPU fallback: CPU tensors, missing library or a non-Blackwell device raise.
"""
import inspect
import json
import math
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from .attention import JointTransformerBlock
from .engine import Engine

SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
CONFIG_NAME = "config.json"


class FrozenConfig(dict):
    """dict with attribute access, like diffusers' FrozenDict (callers use both `.config.x` and `.config["x"]`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("config is frozen")


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


def _sincos_1d(dim, pos):
    omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed_2d(embed_dim, grid_size, base_size, interpolation_scale=1.0):
    """diffusers get_2d_sincos_pos_embed as used by PatchEmbed(pos_embed_max_size=...) (transformer_sd3.py:87-94)."""
    g = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(g, g), axis=0).reshape([2, 1, grid_size, grid_size])   # w first
    emb = np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=1)
    return torch.from_numpy(emb).float().unsqueeze(0)


class PatchEmbed(nn.Module):
    def __init__(self, height, width, patch_size, in_channels, embed_dim, pos_embed_max_size):
        super().__init__()
        self.patch_size = patch_size
        self.pos_embed_max_size = pos_embed_max_size
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=True)
        pe = sincos_pos_embed_2d(embed_dim, pos_embed_max_size, base_size=height // patch_size)
        self.register_buffer("pos_embed", pe, persistent=True)


class _TwoLinear(nn.Module):
    def __init__(self, d_in, d):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, d, bias=True)
        self.linear_2 = nn.Linear(d, d, bias=True)


class CombinedTimestepTextProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.timestep_embedder = _TwoLinear(256, embedding_dim)
        self.text_embedder = _TwoLinear(pooled_projection_dim, embedding_dim)


class _AdaNormOut(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(dim, 2 * dim, bias=True)


class SD3Transformer2DModel(nn.Module):
    _supports_gradient_checkpointing = True
    config_name = CONFIG_NAME
    _ocr_branch = False   # True in SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR (three image streams)

    def __init__(
        self,
        sample_size: int = 128,
        patch_size: int = 2,
        in_channels: int = 16,
        num_layers: int = 18,
        attention_head_dim: int = 64,
        num_attention_heads: int = 18,
        joint_attention_dim: int = 4096,
        caption_projection_dim: int = 1152,
        pooled_projection_dim: int = 2048,
        out_channels: int = 16,
        pos_embed_max_size: int = 96,
        dual_attention_layers: Tuple[int, ...] = (),
        qk_norm: Optional[str] = None,
    ):
        super().__init__()
        cfg = dict(sample_size=sample_size, patch_size=patch_size, in_channels=in_channels, num_layers=num_layers,
                   attention_head_dim=attention_head_dim, num_attention_heads=num_attention_heads,
                   joint_attention_dim=joint_attention_dim, caption_projection_dim=caption_projection_dim,
                   pooled_projection_dim=pooled_projection_dim, out_channels=out_channels,
                   pos_embed_max_size=pos_embed_max_size, dual_attention_layers=tuple(dual_attention_layers),
                   qk_norm=qk_norm, _class_name=type(self).__name__)
        object.__setattr__(self, "_config", FrozenConfig(cfg))
        self.out_channels = out_channels if out_channels is not None else in_channels
        self.inner_dim = num_attention_heads * attention_head_dim

        self.pos_embed = PatchEmbed(sample_size, sample_size, patch_size, in_channels, self.inner_dim, pos_embed_max_size)
        self.time_text_embed = CombinedTimestepTextProjEmbeddings(self.inner_dim, pooled_projection_dim)
        self.context_embedder = nn.Linear(joint_attention_dim, caption_projection_dim)
        self.transformer_blocks = nn.ModuleList([
            JointTransformerBlock(dim=self.inner_dim, num_attention_heads=num_attention_heads,
                                  attention_head_dim=attention_head_dim, context_pre_only=i == num_layers - 1,
                                  qk_norm=qk_norm, use_dual_attention=i in dual_attention_layers,
                                  ocr_branch=self._ocr_branch)
            for i in range(num_layers)])
        self.norm_out = _AdaNormOut(self.inner_dim)
        self.proj_out = nn.Linear(self.inner_dim, patch_size * patch_size * self.out_channels, bias=True)
        self.gradient_checkpointing = False
        self._engine: Optional[Engine] = None

    # ------------------------------------------------------------------ config / io surface used by the reference
    @property
    def config(self) -> FrozenConfig:
        return self._config

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        return cls(**{k: v for k, v in dict(config).items() if k in sig})

    def save_pretrained(self, save_directory, safe_serialization=True, **kwargs):
        """Writes `config.json` + `diffusion_pytorch_model.safetensors`, the pair `from_pretrained_local` reads
        (transformer_sd3.py:417-427; used by the training checkpoint hook, initialize.py:136-137)."""
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()}
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            json.dump(cfg, f, indent=2)
        save_file({k: v.contiguous() for k, v in self.state_dict().items()},
                  os.path.join(save_directory, SAFETENSORS_WEIGHTS_NAME))

    @classmethod
    def from_pretrained_local(cls, pretrained_model_path, subfolder=None, accelerator=None, cfg=None, **kwargs):
        """transformer_sd3.py:412-458: config.json -> from_config; safetensors -> load_state_dict(strict=False);
        missing/unexpected keys are printed and (when accelerator+cfg are given) written to dit4sr_override.txt."""
        import safetensors.torch
        if subfolder is not None:
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        config_file = os.path.join(pretrained_model_path, CONFIG_NAME)
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file, "r") as f:
            config = json.load(f)
        model = cls.from_config(config)
        model_file = os.path.join(pretrained_model_path, SAFETENSORS_WEIGHTS_NAME)
        if not os.path.isfile(model_file):
            raise RuntimeError(f"{model_file} does not exist")
        state_dict = safetensors.torch.load_file(model_file, device="cpu")
        load_result = model.load_state_dict(state_dict, strict=False)
        print("\n===== Missing Keys =====")
        for k in load_result.missing_keys:
            print(f"  - {k}")
        print("\n===== Unexpected Keys =====")
        for k in load_result.unexpected_keys:
            print(f"  - {k}")
        if accelerator is not None and cfg is not None and accelerator.is_main_process:
            txt_file = f"{cfg.save.output_dir}/{cfg.exp_name}/dit4sr_override.txt"
            os.makedirs(os.path.dirname(txt_file), exist_ok=True)
            with open(txt_file, "w") as f:
                f.write("\n===== Missing Keys =====\n")
                for k in load_result.missing_keys:
                    f.write(f"  - {k}\n")
                f.write("\n===== Unexpected Keys =====\n")
                for k in load_result.unexpected_keys:
                    f.write(f"  - {k}\n")
        print("Successfully loaded Transformer weights!")
        return model

    # unused-by-callers API kept as stubs
    def enable_forward_chunking(self, chunk_size: Optional[int] = None, dim: int = 0) -> None:
        if dim not in [0, 1]:
            raise ValueError(f"Make sure to set `dim` to either 0 or 1, not {dim}")
        # the fused MLP kernels never materialise more than one [B,2N,F] buffer; chunking is a no-op here

    def disable_forward_chunking(self):
        pass

    @property
    def attn_processors(self) -> Dict[str, Any]:
        return {}

    def set_attn_processor(self, processor):
        raise NotImplementedError("dit4sr_b200 runs a fixed fused CUDA attention; attention processors cannot be swapped")

    def fuse_qkv_projections(self):
        pass  # the engine always runs the fused q|k|v projection

    def unfuse_qkv_projections(self):
        pass

    def _set_gradient_checkpointing(self, module, value=False):
        if hasattr(module, "gradient_checkpointing"):
            module.gradient_checkpointing = value

    # ------------------------------------------------------------------ engine management
    def invalidate_engine(self):
        """Call after changing parameters in place (the engine keeps packed bf16 copies of the weights)."""
        self._engine = None

    def _apply(self, fn, *args, **kwargs):
        self._engine = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._engine = None
        return super().load_state_dict(*args, **kwargs)

    def engine(self, fold_lr_chains: Optional[bool] = None, sequence_parallel="keep",
               precise: Optional[bool] = None) -> Engine:
        """The CUDA engine (built lazily). `sequence_parallel`: a `dit4sr_b200.parallel.SequenceParallel` to shard ONE
        image over the ranks of its process group (Ulysses), None to turn that off, "keep" to leave it as it is.
        `precise`: force the fp32-accurate mode on/off (default: on iff the parameters are fp32, see Engine)."""
        e = self._engine
        sp = (e.sp if e is not None else None) if isinstance(sequence_parallel, str) else sequence_parallel
        if (e is None or (fold_lr_chains is not None and e.fold_lr_chains != fold_lr_chains) or e.sp is not sp
                or (precise is not None and e.precise != precise)):
            fold = fold_lr_chains if fold_lr_chains is not None else (e.fold_lr_chains if e is not None else None)
            prec = precise if precise is not None else (e.precise if e is not None else None)
            self._engine = Engine(self, fold_lr_chains=fold, sp=sp, precise=prec)
        return self._engine

    # ------------------------------------------------------------------ forward
    def forward(
        self,
        hidden_states: torch.FloatTensor,
        controlnet_image: torch.FloatTensor,
        encoder_hidden_states: torch.FloatTensor = None,
        pooled_projections: torch.FloatTensor = None,
        timestep: torch.LongTensor = None,
        block_controlnet_hidden_states: List = None,
        joint_attention_kwargs: Optional[Dict[str, Any]] = None,
        return_dict: bool = True,
        skip_layers: Optional[List[int]] = None,
        cfg=None,
    ) -> Union[Tuple, Transformer2DModelOutput]:
        if not hidden_states.is_cuda:
            raise RuntimeError("dit4sr_b200.SD3Transformer2DModel runs on a CUDA (sm_100a) device only; "
                               "there is no CPU path (use the oracle under oracle/ for CPU checks)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("dit4sr_b200 implements the inference (denoise) path only; wrap the call in "
                               "torch.no_grad() or call model.requires_grad_(False)")
        if block_controlnet_hidden_states is not None or skip_layers:
            raise NotImplementedError("block_controlnet_hidden_states / skip_layers are never set by DiT4SR's callers")
        # Q7: the reference dereferences cfg.train.model unconditionally; cfg=None here means "no ts_module".
        # (the 3-stream model extracts features for cfg.train.transformer.feat_extract_layer unconditionally and always
        #  returns the (output, trans_blk_outs) pair: transformer_sd3_ocrbranch_ocr2hq2ocr.py:344,371-374,407-408)
        want_feats = cfg is not None and (self._ocr_branch or "ts_module" in cfg.train.model)
        trans_blk_outs = []
        in_dtype = hidden_states.dtype
        features = None
        if want_feats:
            layers = tuple(sorted(set(int(i) for i in cfg.train.transformer.feat_extract_layer)))
            mode = cfg.train.transformer.feat_extract
            # fixed 1024-token slices exactly as written (attention.py:363-374: the reference assumes 64x64 latents)
            slices = {"hq_feat": (0, 1), "lq_feat": (1024, 1), "hqlq_feat": (0, 2)}
            if self._ocr_branch:
                slices["ocr_feat"] = (2048, 1)
            sl = slices.get(mode)
            if sl is not None and layers:
                if hidden_states.shape[-1] != 64 or hidden_states.shape[-2] != 64:
                    raise ValueError("feature extraction slices 1024-token streams (attention.py:363-374): 64x64 latents only")
                tdt = in_dtype if in_dtype in (torch.float32, torch.bfloat16) else torch.float32
                features = (layers, sl[0], sl[1], tdt)
            else:
                trans_blk_outs = [{} for _ in layers]

        hs, ci = hidden_states, controlnet_image
        if in_dtype not in (torch.float32, torch.bfloat16):
            hs, ci = hs.float(), ci.float()
        eng = self.engine()
        if features is None:
            out = eng.step(hs, ci, encoder_hidden_states, pooled_projections, timestep)
        else:
            # the feature kernels run inside the captured graph; each dict carries the reference's token-major slice
            # (`extract_feat`) and the fp32 NCHW tensor the pipeline builds from it (`extract_feat_nchw_f32`,
            # pipelines/pipeline_dit4sr.py:1214-1245: dit4sr_b200.unpatchify_features returns it without recomputing)
            out, feats = eng.step(hs, ci, encoder_hidden_states, pooled_projections, timestep, features=features)
            for d in feats:
                d["extract_feat"] = d["extract_feat"].to(in_dtype)
            trans_blk_outs = feats
        output = out.to(in_dtype)
        if not return_dict:
            if want_feats or self._ocr_branch:
                return (output, trans_blk_outs)
            return (output,)
        return Transformer2DModelOutput(sample=output)
