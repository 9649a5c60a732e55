"""Import the reference's own hot-path modules from /root/reference (TEST INFRASTRUCTURE).

Only usable in the authoring container (the GPU box has no /root/reference);
nothing under `tests/ -m gpu`, `smoke()` or `bench.py` may call this.  It is
used by `tests/golden/make_golden.py` to generate the committed fixtures and
by the optional `tests/test_reference_live.py` (skipped when the tree is absent).

`import revisionllm` fails under the installed transformers 5.5.0
(`vtimellm_llama.py:8-9` imports `SampleDecoderOnlyOutput` & co., removed after
4.4x) and `inference.py` needs `clip`/`easydict`/`decord`/`peft`, all absent.
So: register empty namespace packages so the package `__init__`s never run,
alias the three removed names, and load the individual files by path
(SURVEY.md section 8c).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("RVL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "revisionllm"))


def _ns(name: str, path: str):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def _load(name: str, path: str):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference modules that run on CPU here."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    pkg = os.path.join(REF_ROOT, "revisionllm")
    _ns("revisionllm", pkg)
    _ns("revisionllm.model", os.path.join(pkg, "model"))
    _ns("revisionllm.model.adapter", os.path.join(pkg, "model", "adapter"))
    _ns("revisionllm.eval", os.path.join(pkg, "eval"))
    _ns("revisionllm.uncertainty", os.path.join(pkg, "uncertainty"))

    import transformers.generation as tg
    import transformers.generation.utils as tgu
    for old, new in (("SampleDecoderOnlyOutput", "GenerateDecoderOnlyOutput"),
                     ("SampleEncoderDecoderOutput", "GenerateEncoderDecoderOutput")):
        if not hasattr(tg, old):
            setattr(tg, old, getattr(tgu, new))
    if not hasattr(tg, "validate_stopping_criteria"):
        tg.validate_stopping_criteria = lambda sc, ml: sc
    if not hasattr(tgu, "SampleOutput"):
        tgu.SampleOutput = tgu.GenerateNonBeamOutput

    if "decord" not in sys.modules:          # mm_utils.py:6-8 imports decord at module scope (demo-only use)
        try:
            import decord  # noqa: F401
        except Exception:
            fake = types.ModuleType("decord")
            fake.gpu = fake.cpu = lambda *a, **k: None
            fake.VideoReader = object
            sys.modules["decord"] = fake

    out = types.SimpleNamespace()
    out.constants = _load("revisionllm.constants", os.path.join(pkg, "constants.py"))
    out.conversation = _load("revisionllm.conversation", os.path.join(pkg, "conversation.py"))
    out.transformer = _load("revisionllm.model.adapter.transformer", os.path.join(pkg, "model", "adapter", "transformer.py"))
    out.tensor_utils = _load("revisionllm.model.adapter.tensor_utils", os.path.join(pkg, "model", "adapter", "tensor_utils.py"))
    out.arch = _load("revisionllm.model.vtimellm_arch", os.path.join(pkg, "model", "vtimellm_arch.py"))
    out.llama = _load("revisionllm.model.vtimellm_llama", os.path.join(pkg, "model", "vtimellm_llama.py"))
    out.mm_utils = _load("revisionllm.mm_utils", os.path.join(pkg, "mm_utils.py"))
    out.similarity = _load("revisionllm.eval.similarity", os.path.join(pkg, "eval", "similarity.py"))
    out.entropy = _load("revisionllm.uncertainty.funs_get_feature_X", os.path.join(pkg, "uncertainty", "funs_get_feature_X.py"))
    return out


def build_reference_model(ref, cfg, weights, clip_weights=None):
    """Instantiate the reference's `VTimeLLMLlamaForCausalLM` (fp32, eager
    attention) and load synthetic weights.  `clip_weights` switches the adapter
    to the stage-2 `ClipEncoder` (hierarchy / clip_adapter_text)."""
    import torch
    conf = ref.llama.VTimeLLMConfig(
        hidden_size=cfg.hidden, intermediate_size=cfg.intermediate, num_hidden_layers=cfg.n_layers,
        num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_heads, vocab_size=cfg.vocab,
        rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_pos,
        attn_implementation="eager", pretraining_tp=1, tie_word_embeddings=False,
    )
    conf.pretraining_tp = 1
    conf._attn_implementation = "eager"
    model = ref.llama.VTimeLLMLlamaForCausalLM(conf)
    args = types.SimpleNamespace(
        clip_adapter=clip_weights is not None, cross_attn=False, pretrain_clip_adapter=None,
        pretrain_mm_mlp_adapter=None, clip_adapter_text=clip_weights is not None,
        clip_adapter_feature="cls", hierarchy=clip_weights is not None, adapter_input_dim=cfg.adapter_dim,
    )
    model.get_model().initialize_vision_modules(args)
    sd = {k: v.float() for k, v in weights.items() if not k.startswith("model.mm_projector.")}
    if clip_weights is None:
        sd["model.mm_projector.weight"] = weights["model.mm_projector.weight"].float()
        sd["model.mm_projector.bias"] = weights["model.mm_projector.bias"].float()
    else:
        for k, v in clip_weights.items():
            sd["model.mm_projector." + k] = v.float()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "rotary_emb" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return model.float().eval()
