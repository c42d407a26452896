"""dit4sr_b200: B200 (sm_100a) implementation of DiT4SR's denoising hot path behind the reference's module API."""
from .transformer_sd3 import SD3Transformer2DModel, Transformer2DModelOutput  # noqa: F401
from .transformer_sd3_ocrbranch_ocr2hq2ocr import SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR  # noqa: F401
from .features import unpatchify_features  # noqa: F401
from .vae import AutoencoderKL  # noqa: F401
from .attention import JointTransformerBlock, AttentionZero, FeedForwardControl, FeedForward  # noqa: F401

__all__ = ["SD3Transformer2DModel", "SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR", "Transformer2DModelOutput", "JointTransformerBlock", "AttentionZero",
           "FeedForwardControl", "FeedForward", "unpatchify_features", "AutoencoderKL"]
