"""Drop-in for `model_dit4sr/transformer_sd3_ocrbranch_ocr2hq2ocr.py::SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR`:
the 3-stream [HQ | LR | OCR] variant (BASELINE config #5). Same CUDA engine, three image streams:

  * the OCR stream is a second `pos_embed(controlnet_image)` (reference :334-337);
  * `attn` attends over [HQ | LR | OCR | text] with `to_*_ocr` chains and a live `norm_k` on the OCR keys
    (attention_ocrbranch_ocr2hq2ocr.py:520-664); `attn2` remains the 2-stream processor and therefore splits the 3N image
    tokens at 1.5N (quirk Q6);
  * the MLP has three depthwise convs: LR->HQ, OCR->HQ, then HQ(updated)->OCR (:164-211);
  * forward always returns `(output, trans_blk_outs)` when `return_dict=False` (:407-408).
"""
import json
import os

from .transformer_sd3 import CONFIG_NAME, SAFETENSORS_WEIGHTS_NAME, SD3Transformer2DModel


class SD3Transformer2DModel_OCRBranch_OCR2HQ2OCR(SD3Transformer2DModel):
    _ocr_branch = True

    @classmethod
    def from_pretrained_local(cls, pretrained_model_path, subfolder=None, accelerator=None, cfg=None, **kwargs):
        """transformer_sd3_ocrbranch_ocr2hq2ocr.py:414-525. A plain DiT4SR checkpoint has no `*_ocr` tensors; when
        `cfg.train.transformer.ocr_branch_init` is 'hq_branch' / 'lq_branch' they are initialised from the HQ
        (`to_q`, `to_out.0`, ...) or LR (`to_q_control`, `to_out_control`, ...) projections of the same block; the two
        new convs stay zero."""
        import safetensors.torch
        if subfolder is not None:
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        config_file = os.path.join(pretrained_model_path, CONFIG_NAME)
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file, "r") as f:
            model = cls.from_config(json.load(f))
        model_file = os.path.join(pretrained_model_path, SAFETENSORS_WEIGHTS_NAME)
        if not os.path.isfile(model_file):
            raise RuntimeError(f"{model_file} does not exist")
        state_dict = safetensors.torch.load_file(model_file, device="cpu")

        init = None if cfg is None else getattr(cfg.train.transformer, "ocr_branch_init", None)
        log = []
        if init is not None:
            if init not in ("hq_branch", "lq_branch"):
                raise ValueError(f"unknown ocr_branch_init: {init}")
            missing = [k for k in model.state_dict().keys()
                       if k not in state_dict and "hidden_to_ocr_conv" not in k and "ocr_to_hidden_conv" not in k]
            for key in missing:
                parts = key.split(".")          # transformer_blocks.<i>.attn.to_q_ocr.weight
                name = parts[3]
                if not name.endswith("_ocr"):
                    raise ValueError(f"[OCR weight mapping] no source rule for '{key}'")
                if init == "lq_branch":
                    src = name.replace("_ocr", "_control")
                else:
                    src = name.replace("_ocr", ".0" if name == "to_out_ocr" else "")
                src_key = ".".join(parts[:3] + [src] + parts[4:])
                if src_key not in state_dict:
                    raise ValueError(f"[OCR weight mapping] source '{src_key}' for '{key}' not in the checkpoint")
                state_dict[key] = state_dict[src_key].clone()
                log.append(f"Copied:  {key:<50}  <-  {src_key}")
        result = model.load_state_dict(state_dict, strict=False)
        print("\n===== Missing Keys =====")
        for k in result.missing_keys:
            print(f"  - {k}")
        print("\n===== Unexpected Keys =====")
        for k in result.unexpected_keys:
            print(f"  - {k}")
        if accelerator is not None and cfg is not None and accelerator.is_main_process:
            txt_file = f"{cfg.save.output_dir}/{cfg.exp_name}/dit4sr_override.txt"
            os.makedirs(os.path.dirname(txt_file), exist_ok=True)
            with open(txt_file, "w") as f:
                f.write("\n".join(log) + "\n\n===== Missing Keys =====\n")
                f.writelines(f"  - {k}\n" for k in result.missing_keys)
                f.write("\n===== Unexpected Keys =====\n")
                f.writelines(f"  - {k}\n" for k in result.unexpected_keys)
        print("Successfully loaded Transformer weights!")
        return model
