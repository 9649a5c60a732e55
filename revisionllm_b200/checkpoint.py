"""Real-weight loading: HF checkpoint + mm_projector.bin + PEFT LoRA merge -> the
HF-named bf16 state dict `Engine.bind_state_dict` takes.

Mirrors /root/reference/revisionllm/model/builder.py:
  * `load_pretrained_model` (:21-67): base Vicuna weights (`from_pretrained`), `initialize_vision_modules` reading
    `mm_projector.bin` (vtimellm_arch.py:12-73), then stage-2 / stage-3 LoRA adapters loaded and merged one after the other
    (`PeftModel.from_pretrained(...).merge_and_unload()`, :47-60);
  * `load_lora` (:9-19): `non_lora_trainables.bin` with its `base_model.` / `model.` key prefixes stripped.
PEFT's merge rule for a Linear is W <- W + (lora_alpha / r) * (lora_B @ lora_A) (peft/tuners/lora/layer.py `get_delta_weight`;
peft is a requirement of the reference, `peft>=0.4.0`, and absent from this image).  Here lora_B . lora_A is one tcgen05
GEMM per target module (K = r) with fp32 accumulation, added to the weight in fp32 and rounded to bf16 once.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterable, Optional

import torch

from ._cabi import GEMM_OUT_F32, RvlError


def load_state_dict_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def load_hf_checkpoint(model_dir: str) -> Dict[str, torch.Tensor]:
    """All shards of a HF checkpoint directory (model.safetensors[.index.json] or pytorch_model.bin[.index.json])."""
    for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
        p = os.path.join(model_dir, index)
        if os.path.exists(p):
            files = sorted(set(json.load(open(p))["weight_map"].values()))
            sd: Dict[str, torch.Tensor] = {}
            for f in files:
                sd.update(load_state_dict_file(os.path.join(model_dir, f)))
            return sd
    for single in ("model.safetensors", "pytorch_model.bin"):
        p = os.path.join(model_dir, single)
        if os.path.exists(p):
            return load_state_dict_file(p)
    raise RvlError(f"no HF checkpoint found in {model_dir}")


def strip_lora_prefixes(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Key clean-up of builder.py:12-15 for non_lora_trainables.bin."""
    sd = {(k[11:] if k.startswith("base_model.") else k): v for k, v in sd.items()}
    if any(k.startswith("model.model.") for k in sd):
        sd = {(k[6:] if k.startswith("model.") else k): v for k, v in sd.items()}
    return sd


_LORA_KEY = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+)\.lora_(?P<ab>[AB])(?:\.default)?\.weight$")


def merge_lora(engine, state_dict: Dict[str, torch.Tensor], lora_sd: Dict[str, torch.Tensor], lora_alpha: float, r: Optional[int] = None,
               ) -> Dict[str, torch.Tensor]:
    """state_dict[<module>.weight] += (lora_alpha / r) * lora_B @ lora_A for every adapter pair in `lora_sd`
    (keys `base_model.model.<module>.lora_A[.default].weight` [r, in], `...lora_B...` [out, r]).  Returns the same dict with
    the merged tensors replaced (bf16, on the engine's device)."""
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, v in lora_sd.items():
        m = _LORA_KEY.match(k)
        if m:
            pairs.setdefault(m.group("mod"), {})[m.group("ab")] = v
    dev = engine.device
    for mod, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise RvlError(f"incomplete LoRA pair for {mod}")
        name = mod + ".weight"
        if name not in state_dict:
            raise RvlError(f"LoRA target {name} is not in the base state dict")
        A, B = ab["A"], ab["B"]
        rank = A.shape[0] if r is None else r
        scale = float(lora_alpha) / float(rank)
        W = state_dict[name].to(dev)
        if A.shape[1] != W.shape[1] or B.shape[0] != W.shape[0] or B.shape[1] != A.shape[0]:
            raise RvlError(f"LoRA shapes do not match {name}: A {tuple(A.shape)}, B {tuple(B.shape)}, W {tuple(W.shape)}")
        k_pad = (-A.shape[0]) % 8                                   # the GEMM wants K % 8 == 0
        Bs = (B.to(dev, torch.float32) * scale).to(torch.bfloat16)
        At = A.to(dev, torch.bfloat16).t().contiguous()             # [in, r]: the GEMM takes both operands K-major
        if k_pad:
            Bs = torch.nn.functional.pad(Bs, (0, k_pad))
            At = torch.nn.functional.pad(At, (0, k_pad))
        delta = engine.gemm(Bs.contiguous(), At.contiguous(), out_mode=GEMM_OUT_F32)    # [out, in] fp32 = scale * B @ A
        state_dict[name] = (W.float() + delta).to(torch.bfloat16)
    return state_dict


def build_state_dict(engine, model_base: str, pretrain_mm_mlp_adapter: Optional[str] = None, stages: Iterable[str] = (),
                     lora_alpha: float = 128.0, lora_r: int = 64) -> Dict[str, torch.Tensor]:
    """`load_pretrained_model` at the tensor level: base checkpoint, projector weights, then every stage directory's
    `non_lora_trainables.bin` and `adapter_model.{safetensors,bin}` merged in order (scripts/mad/stage1_dense.sh:55-56:
    r = 64, alpha = 128).  Bind the result with `Engine.bind_state_dict`."""
    sd = {k: v.to(torch.bfloat16) for k, v in load_hf_checkpoint(model_base).items()}
    if pretrain_mm_mlp_adapter is not None:
        proj = torch.load(pretrain_mm_mlp_adapter, map_location="cpu", weights_only=True)
        for k, v in proj.items():                                   # initialize_vision_modules keeps the part after 'mm_projector.'
            if "mm_projector" in k:
                sd["model.mm_projector." + k.split("mm_projector.")[1]] = v.to(torch.bfloat16)
    for stage in stages:
        extra = os.path.join(stage, "non_lora_trainables.bin")
        if os.path.exists(extra):
            sd.update({k: v.to(torch.bfloat16) for k, v in strip_lora_prefixes(torch.load(extra, map_location="cpu", weights_only=True)).items()})
        alpha, rank = lora_alpha, lora_r
        cfg = os.path.join(stage, "adapter_config.json")
        if os.path.exists(cfg):
            c = json.load(open(cfg))
            alpha, rank = c.get("lora_alpha", alpha), c.get("r", rank)
        for f in ("adapter_model.safetensors", "adapter_model.bin"):
            p = os.path.join(stage, f)
            if os.path.exists(p):
                merge_lora(engine, sd, load_state_dict_file(p), alpha, rank)
                break
    return sd
