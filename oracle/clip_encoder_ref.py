"""Stage-2 adapter `ClipEncoder` restated in explicit fp32 math (TEST INFRASTRUCTURE).

Follows /root/reference/revisionllm/model/adapter/transformer.py:
  * PositionEmbeddingSine.forward           :35-57  (normalize=True, scale=2*pi, temperature=1e4)
  * ClipEncoder.forward                     :94-145 (clip_adapter_text=True, hierarchy=True -> CLS row)
  * T2V_TransformerEncoderLayer.forward_post :271-305
  * TransformerEncoderLayer.forward_post     :210-223
and the hierarchy branch of the splice that calls it,
/root/reference/revisionllm/model/vtimellm_arch.py:114-121.
`nn.MultiheadAttention` (torch) is restated as in_proj -> per-head scaled
dot-product softmax (key padding -> -inf) -> out_proj.  Dropout is inactive at
inference.  Parameter names are the reference module's state_dict keys below
the prefix `model.mm_projector.`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

D_MODEL = 768
N_HEAD = 8
N_LAYERS = 2
FFN = 2048


def sine_pos(n_frames: int, d: int = D_MODEL, temperature: float = 10000.0) -> torch.Tensor:
    """:45-55 with an all-ones mask: x_embed = (1..T)/(T + 1e-6) * 2*pi."""
    x_embed = torch.arange(1, n_frames + 1, dtype=torch.float32)
    x_embed = x_embed / (x_embed[-1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(d, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / d)
    pos = x_embed[:, None] / dim_t
    return torch.stack((pos[:, 0::2].sin(), pos[:, 1::2].cos()), dim=2).flatten(1)  # [T, d]


def _mha(p: Dict[str, torch.Tensor], pre: str, q, k, v, key_pad: Optional[torch.Tensor]):
    """q [B,Lq,d], k/v [B,Lk,d], key_pad [B,Lk] True = ignore."""
    W, b = p[pre + "in_proj_weight"].float(), p[pre + "in_proj_bias"].float()
    d = q.shape[-1]
    hd = d // N_HEAD
    qp = F.linear(q, W[:d], b[:d])
    kp = F.linear(k, W[d:2 * d], b[d:2 * d])
    vp = F.linear(v, W[2 * d:], b[2 * d:])
    B, Lq, _ = qp.shape
    Lk = kp.shape[1]
    qp = qp.view(B, Lq, N_HEAD, hd).transpose(1, 2)
    kp = kp.view(B, Lk, N_HEAD, hd).transpose(1, 2)
    vp = vp.view(B, Lk, N_HEAD, hd).transpose(1, 2)
    s = torch.matmul(qp, kp.transpose(2, 3)) / math.sqrt(hd)
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(a, vp).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, p[pre + "out_proj.weight"].float(), p[pre + "out_proj.bias"].float())


def _ln(p, pre, x):
    return F.layer_norm(x, (x.shape[-1],), p[pre + "weight"].float(), p[pre + "bias"].float(), 1e-5)


def _ffn(p, pre, x):
    h = F.relu(F.linear(x, p[pre + "linear1.weight"].float(), p[pre + "linear1.bias"].float()))
    return F.linear(h, p[pre + "linear2.weight"].float(), p[pre + "linear2.bias"].float())


def clip_encoder_cls(
    p: Dict[str, torch.Tensor],            # keys below 'model.mm_projector.'
    frames: torch.Tensor,                  # [V, T, 768]
    text: torch.Tensor,                    # [V, Lq, 768]
    text_mask: torch.Tensor,               # [V, Lq] 1 = valid
    return_memory: bool = False,
) -> torch.Tensor:
    """ClipEncoder.forward -> mm_projector(memory[0]) : [V, hidden]."""
    frames, text = frames.float(), text.float()
    V, T, d = frames.shape
    pos = sine_pos(T, d)[None].expand(V, T, d)
    x = frames
    # --- 2x text->video cross attention layers (:117-124, :271-305)
    key_pad = ~text_mask.bool()
    for i in range(N_LAYERS):
        pre = f"t2v_encoder.layers.{i}."
        q = x + pos
        k = text                       # pos_txt = 0
        src2 = x + _mha(p, pre + "self_attn.", q, k, text, key_pad)
        src3 = _ffn(p, pre, _ln(p, pre + "norm1.", src2))
        x = _ln(p, pre + "norm2.", src2 + src3)
    # --- prepend global token, 2x post-norm self-attention layers (:133, :210-223)
    g = p["global_rep_token"].float().view(1, 1, d).expand(V, 1, d)
    gp = p["global_rep_pos"].float().view(1, 1, d).expand(V, 1, d)
    x = torch.cat((g, x), dim=1)
    pe = torch.cat((gp, pos), dim=1)
    for i in range(N_LAYERS):
        pre = f"encoder.layers.{i}."
        qk = x + pe
        x = _ln(p, pre + "norm1.", x + _mha(p, pre + "self_attn.", qk, qk, x, None))
        x = _ln(p, pre + "norm2.", x + _ffn(p, pre, x))
    if return_memory:
        return x
    return F.linear(x[:, 0], p["mm_projector.weight"].float(), p["mm_projector.bias"].float())


def hierarchy_features(p, images: torch.Tensor, query_feats, ) -> torch.Tensor:
    """vtimellm_arch.py:114-121: images [b, v, t, d] -> [(b v), t, d]; the
    query tokens/mask are repeated per segment; result [b, v, hidden]."""
    b, v, t, d = images.shape
    q_tok, q_mask = query_feats
    qt = q_tok[:, None].repeat(1, v, 1, 1).reshape(b * v, q_tok.shape[1], q_tok.shape[2])
    qm = q_mask[:, None].repeat(1, v, 1).reshape(b * v, q_mask.shape[1])
    out = clip_encoder_cls(p, images.reshape(b * v, t, d), qt, qm)
    return out.view(b, v, -1)
