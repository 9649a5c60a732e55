"""fp32 CPU restatement of the Llama decoder the reference runs (TEST INFRASTRUCTURE).

The reference never implements the decoder itself: `VTimeLLMLlamaForCausalLM`
(/root/reference/revisionllm/model/vtimellm_llama.py:23-36) subclasses
`transformers.LlamaForCausalLM` (pinned transformers==4.41.2,
/root/reference/requirements.txt:10) and reaches it through `super().forward`
(vtimellm_llama.py:79-90).  The published algorithm restated here is the one in
`transformers/models/llama/modeling_llama.py` (LlamaRMSNorm, rotate_half RoPE,
eager attention with fp32 softmax, SwiGLU MLP, final norm, lm_head, `.float()`
logits) - identical between 4.41.2 and the 5.5.0 installed in the authoring
container up to refactoring.  `tests/golden/make_golden.py` pins this file
against that installed implementation driven by the reference's own
`forward`/splice code.

All arithmetic is fp32.  Weights are whatever tensors the caller passes (tests
pass bf16-representable values up-cast to fp32, so the oracle and the bf16 CUDA
path see the same parameters).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class LlamaShape:
    """Vicuna-7B-v1.5 = Llama-2-7B shape by default."""
    hidden: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    head_dim: int = 128
    intermediate: int = 11008
    vocab: int = 32000
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    adapter_dim: int = 768


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm: x * rsqrt(mean(x^2) + eps) * w, all in fp32."""
    x = x.float()
    var = x.pow(2).mean(-1, keepdim=True)
    return w.float() * (x * torch.rsqrt(var + eps))


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """LlamaRotaryEmbedding (default rope): inv_freq_i = theta^(-2i/d), emb = cat(freqs, freqs)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = positions.float()[..., None] * inv_freq  # [..., d/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


class KVCache:
    """Per-layer list of (k, v) tensors shaped [B, heads, L, d] (DynamicCache equivalent)."""

    def __init__(self, n_layers: int):
        self.k: List[Optional[torch.Tensor]] = [None] * n_layers
        self.v: List[Optional[torch.Tensor]] = [None] * n_layers

    def length(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[2]

    def append(self, layer: int, k: torch.Tensor, v: torch.Tensor):
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat((self.k[layer], k), dim=2)
            self.v[layer] = torch.cat((self.v[layer], v), dim=2)
        return self.k[layer], self.v[layer]


def decoder_stack(
    w: Dict[str, torch.Tensor],
    shape: LlamaShape,
    inputs_embeds: torch.Tensor,          # [B, L, hidden] fp32
    cache: Optional[KVCache] = None,
    collect: Optional[dict] = None,
) -> torch.Tensor:
    """LlamaModel.forward for a batch of equal-length, unpadded sequences.

    Positions continue from the cache length (the reference's decode fix-up,
    /root/reference/revisionllm/model/vtimellm_arch.py:88-100, recovers exactly
    that for right-padded batches: position = number of valid tokens so far).
    Returns the final-norm hidden states [B, L, hidden].
    """
    B, L, H = inputs_embeds.shape
    nh, d = shape.n_heads, shape.head_dim
    past = cache.length() if cache is not None else 0
    pos = torch.arange(past, past + L)
    cos, sin = rope_cos_sin(pos, d, shape.rope_theta)        # [L, d]
    cos, sin = cos[None, None], sin[None, None]
    # causal mask over [L, past+L]
    qi = torch.arange(L)[:, None] + past
    kj = torch.arange(past + L)[None, :]
    mask = torch.where(kj <= qi, 0.0, float("-inf"))
    h = inputs_embeds.float()
    for i in range(shape.n_layers):
        p = f"model.layers.{i}."
        x = rmsnorm(h, w[p + "input_layernorm.weight"], shape.rms_eps)
        q = F.linear(x, w[p + "self_attn.q_proj.weight"].float()).view(B, L, nh, d).transpose(1, 2)
        k = F.linear(x, w[p + "self_attn.k_proj.weight"].float()).view(B, L, nh, d).transpose(1, 2)
        v = F.linear(x, w[p + "self_attn.v_proj.weight"].float()).view(B, L, nh, d).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        if cache is not None:
            k, v = cache.append(i, k, v)
        att = torch.matmul(q, k.transpose(2, 3)) * (1.0 / math.sqrt(d)) + mask
        att = torch.softmax(att, dim=-1, dtype=torch.float32)
        o = torch.matmul(att, v).transpose(1, 2).reshape(B, L, nh * d)
        h = h + F.linear(o, w[p + "self_attn.o_proj.weight"].float())
        x = rmsnorm(h, w[p + "post_attention_layernorm.weight"], shape.rms_eps)
        g = F.linear(x, w[p + "mlp.gate_proj.weight"].float())
        u = F.linear(x, w[p + "mlp.up_proj.weight"].float())
        h = h + F.linear(F.silu(g) * u, w[p + "mlp.down_proj.weight"].float())
        if collect is not None:
            collect.setdefault("layer_out", []).append(h.clone())
    return rmsnorm(h, w["model.norm.weight"], shape.rms_eps)


def lm_head(w: Dict[str, torch.Tensor], hidden: torch.Tensor) -> torch.Tensor:
    """`logits = lm_head(hidden).float()` (LlamaForCausalLM.forward)."""
    return F.linear(hidden.float(), w["lm_head.weight"].float())


def embed_tokens(w: Dict[str, torch.Tensor], ids: torch.Tensor) -> torch.Tensor:
    return w["model.embed_tokens.weight"].float()[ids]


def eos_bookkeeping(next_tokens: torch.Tensor, unfinished: torch.Tensor, eos_token_id: int, pad_token_id: int
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """One step of the reference's EOS rule (/root/reference/revisionllm/model/vtimellm_llama.py:340-356): rows that have
    finished emit `pad_token_id`; a row finishes with the step in which it emits EOS.  Returns (tokens appended this step,
    updated unfinished flags).  Pinned by tests/golden/eos_rule.json (the reference's own lines, exec'd)."""
    tokens = next_tokens * unfinished + pad_token_id * (1 - unfinished)
    return tokens, unfinished * (tokens != eos_token_id).long()


def greedy_decode(
    w: Dict[str, torch.Tensor],
    shape: LlamaShape,
    inputs_embeds: torch.Tensor,           # [B, L, hidden] (already spliced)
    max_new_tokens: int,
    eos_token_id: Optional[int] = 2,
    pad_token_id: Optional[int] = None,
    stop_on_eos: bool = True,
) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Greedy (argmax) variant of the generation loop the reference drives.

    Follows the bookkeeping of the reference's copy of HF `sample()`
    (/root/reference/revisionllm/model/vtimellm_llama.py:287-369): last-row
    logits are the step's `scores` (raw, `:321`), finished rows emit
    `pad_token_id` (`:343-347`), a row finishes when it emits EOS (`:352-356`)
    and the loop stops when every row has finished (`:359-362`) or after
    `max_new_tokens`.  The multinomial draw (`:337-338`) is replaced by argmax
    with lowest-index tie-break (BASELINE.json north_star: greedy decode).
    Returns (new_tokens [B, T'], scores list of T' tensors [B, vocab]).
    """
    B = inputs_embeds.shape[0]
    if pad_token_id is None:
        pad_token_id = eos_token_id if eos_token_id is not None else 0
    cache = KVCache(shape.n_layers)
    hidden = decoder_stack(w, shape, inputs_embeds, cache)
    logits = lm_head(w, hidden[:, -1])
    unfinished = torch.ones(B, dtype=torch.long)
    toks, scores = [], []
    for _ in range(max_new_tokens):
        scores.append(logits.clone())
        nxt = torch.argmax(logits, dim=-1)
        if stop_on_eos and eos_token_id is not None:
            nxt, unfinished = eos_bookkeeping(nxt, unfinished, eos_token_id, pad_token_id)
        toks.append(nxt)
        if stop_on_eos and unfinished.max() == 0:
            break
        if len(toks) == max_new_tokens:
            break
        hidden = decoder_stack(w, shape, embed_tokens(w, nxt)[:, None], cache)
        logits = lm_head(w, hidden[:, -1])
    return torch.stack(toks, dim=1), scores


def forward_ragged(
    w: Dict[str, torch.Tensor],
    shape: LlamaShape,
    embeds: Sequence[torch.Tensor],        # list of [L_i, hidden]
    max_new_tokens: int,
    **kw,
) -> Tuple[List[torch.Tensor], List[List[torch.Tensor]]]:
    """Ragged batch = independent sequences.  The reference right-pads them and
    masks the padding (/root/reference/revisionllm/model/vtimellm_arch.py:246-276),
    which in exact arithmetic equals running each sequence alone; a finished
    row keeps stepping (with pad tokens) until all rows finish, which the
    caller can reproduce by passing stop_on_eos=False and truncating."""
    toks, scs = [], []
    for e in embeds:
        t, s = greedy_decode(w, shape, e[None], max_new_tokens, **kw)
        toks.append(t[0])
        scs.append([x[0] for x in s])
    return toks, scs
