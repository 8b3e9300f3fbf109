"""Model specs and loaders -- drop-ins for /root/reference/Continuous/src/flux/util.py:111-157 (``ModelSpec``,
``configs``), :210-223 (``load_flow_model`` / ``load_flow_model2``) and :227-246 (``load_ae``).

Same names, arguments and behaviour: ``load_flow_model2`` returns a random-initialised fp32 ``Flux`` (the caller
casts it to bf16, train_SigLIP_stage1.py:131-132); ``load_ae`` reads the safetensors file named by ``$AE`` when it
is set and otherwise random-initialises on ``device``.  The reference's unused imports (optimum.quanto, cv2,
huggingface_hub) are not reproduced.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from .model import Flux, FluxParams
from .modules.autoencoder import AutoEncoder, AutoEncoderParams


@dataclass
class ModelSpec:
    params: FluxParams
    ae_params: AutoEncoderParams
    ckpt_path: str | None
    ae_path: str | None
    repo_id: str | None
    repo_flow: str | None
    repo_ae: str | None
    repo_id_ae: str | None


def _flux_dev_params() -> FluxParams:  # util.py:131-144 -- the "lightweight denoiser": depth 2 + 4
    return FluxParams(in_channels=64, vec_in_dim=768, context_in_dim=4096, hidden_size=3072, mlp_ratio=4.0,
                      num_heads=24, depth=2, depth_single_blocks=4, axes_dim=[16, 56, 56], theta=10_000, qkv_bias=True,
                      guidance_embed=True)


def _ae_params() -> AutoEncoderParams:  # util.py:146-156
    return AutoEncoderParams(resolution=256, in_channels=3, ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
                             z_channels=16, scale_factor=0.3611, shift_factor=0.1159)


configs = {
    "flux-dev": ModelSpec(repo_id="black-forest-labs/FLUX.1-dev", repo_id_ae="black-forest-labs/FLUX.1-dev",
                          repo_flow="flux1-dev.safetensors", repo_ae="ae.safetensors", ckpt_path=None,
                          params=_flux_dev_params(), ae_path=os.getenv("AE"), ae_params=_ae_params()),
}


def print_load_warning(missing: list[str], unexpected: list[str]) -> None:
    if missing:
        print(f"Got {len(missing)} missing keys:\n\t" + "\n\t".join(missing))
    if unexpected:
        print(f"Got {len(unexpected)} unexpected keys:\n\t" + "\n\t".join(unexpected))


def load_flow_model(name: str, device: str | torch.device = "cuda", hf_download: bool = True) -> Flux:
    return Flux(configs[name].params).to(torch.bfloat16)


def load_flow_model2(name: str, device: str | torch.device = "cuda", hf_download: bool = True) -> Flux:
    print("Random init flux...")
    return Flux(configs[name].params)


def load_ae(name: str, device: str | torch.device = "cuda", hf_download: bool = True) -> AutoEncoder:
    ckpt_path = configs[name].ae_path or os.getenv("AE")
    print("Init AE")
    with torch.device(device):
        ae = AutoEncoder(configs[name].ae_params)
    if ckpt_path is not None:
        from safetensors.torch import load_file as load_sft
        sd = load_sft(ckpt_path, device=str(device))
        missing, unexpected = ae.load_state_dict(sd, strict=False)
        print_load_warning(list(missing), list(unexpected))
    return ae
