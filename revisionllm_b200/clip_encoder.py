"""Stage-2 adapter `ClipEncoder` composed from the C-ABI kernels (tcgen05 GEMM, LayerNorm, MHA-96).

Mirrors /root/reference/revisionllm/model/adapter/transformer.py:
  ClipEncoder.forward :94-145 (clip_adapter_text=True, hierarchy=True -> project the CLS row),
  T2V_TransformerEncoderLayer.forward_post :271-305, TransformerEncoderLayer.forward_post :210-223,
  PositionEmbeddingSine.forward :35-57 (a constant [T, 768] table, built once on the host in fp32),
and the hierarchy branch that calls it, /root/reference/revisionllm/model/vtimellm_arch.py:114-121.

Data layout: the encoder's residual stream is fp32 [V*T, 768] (then [V*(1+T), 768] once the global
token is prepended); GEMM operands are bf16 copies produced by the LayerNorm kernel's extra outputs
(`y_bf16`, `y + pos`), GEMM epilogues add bias / ReLU / the fp32 residual in place.  The text keys and
values are projected once per query, not once per segment (the reference repeats the query V times).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from ._cabi import GEMM_ADD_F32, GEMM_FLAG_RELU, GEMM_OUT_BF16, GEMM_OUT_F32
from .engine import Engine

D, NH, FFN, NL = 768, 8, 2048, 2


def sine_position_table(n: int, d: int = D, temperature: float = 10000.0) -> torch.Tensor:
    """PositionEmbeddingSine(normalize=True) for an all-ones mask: x_embed = (1..n)/(n + 1e-6) * 2*pi,
    sin on even / cos on odd channels (transformer.py:45-55).  fp32 [n, d]."""
    x = torch.arange(1, n + 1, dtype=torch.float32)
    x = x / (x[-1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(d, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / d)
    p = x[:, None] / dim_t
    return torch.stack((p[:, 0::2].sin(), p[:, 1::2].cos()), dim=2).flatten(1)


class ClipEncoder:
    def __init__(self, engine: Engine, state: Dict[str, torch.Tensor]):
        self.eng = engine
        dev = engine.device
        self.p = {k: v.to(dev, torch.bfloat16).contiguous() for k, v in state.items()}
        self.hidden = self.p["mm_projector.weight"].shape[0]
        self._pos_cache: Dict[int, tuple] = {}

    def _pos(self, T: int):
        if T not in self._pos_cache:
            dev = self.eng.device
            pos = sine_position_table(T).to(dev)
            gpos = self.p["global_rep_pos"].float()[None]
            self._pos_cache[T] = (pos.contiguous(), torch.cat([gpos, pos], dim=0).contiguous())
        return self._pos_cache[T]

    def _c_weights(self, T: int):
        """The adapter's parameters as the `rvl_clip_weights` table of the C entry point (pointers into self.p)."""
        from . import _cabi
        pos, pos_g = self._pos(T)
        w = _cabi.rvl_clip_weights()

        def layer(pre):
            names = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                     "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")
            return _cabi.rvl_clip_layer(*[self.p[pre + n].data_ptr() for n in names])
        for i in range(NL):
            w.t2v[i] = layer(f"t2v_encoder.layers.{i}.")
            w.enc[i] = layer(f"encoder.layers.{i}.")
        w.global_token = self.p["global_rep_token"].data_ptr()
        w.pos, w.pos_global = pos.data_ptr(), pos_g.data_ptr()
        w.proj_w, w.proj_b = self.p["mm_projector.weight"].data_ptr(), self.p["mm_projector.bias"].data_ptr()
        w.hidden = self.hidden
        return w

    def __call__(self, frames: torch.Tensor, text: torch.Tensor, text_mask: torch.Tensor,
                 seg_text_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames [V, T, 768] bf16; text [Q, Lq, 768] bf16; text_mask [Q, Lq] (1 = valid);
        seg_text_idx int32 [V]: which text each segment attends to (None: Q == V, one to one).
        Returns the projected CLS rows [V, hidden] bf16.  One call into the C ABI (rvl_clip_encoder); `composed()` is the same
        sequence of kernels driven from here, kept for the unit test that compares the two bit for bit."""
        if not hasattr(self.eng, "clip_encoder"):
            return self.composed(frames, text, text_mask, seg_text_idx)       # engines without the one-call entry (test stand-ins)
        dev = self.eng.device
        V, T, d = frames.shape
        Q, Lq, _ = text.shape
        assert d == D
        if seg_text_idx is None and Q != V:
            raise ValueError("seg_text_idx is required when the number of texts differs from the number of segments")
        out = torch.empty((V, self.hidden), dtype=torch.bfloat16, device=dev)
        return self.eng.clip_encoder(self._c_weights(T), frames.to(dev, torch.bfloat16).contiguous(), text.to(dev, torch.bfloat16).contiguous(),
                                     text_mask.to(dev, torch.float32).contiguous(),
                                     None if seg_text_idx is None else seg_text_idx.to(dev, torch.int32).contiguous(), out)

    def composed(self, frames: torch.Tensor, text: torch.Tensor, text_mask: torch.Tensor,
                 seg_text_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        eng, p = self.eng, self.p
        dev = eng.device
        V, T, d = frames.shape
        Q, Lq, _ = text.shape
        assert d == D
        pos, pos_g = self._pos(T)
        frames = frames.to(dev, torch.bfloat16).contiguous()
        text2d = text.to(dev, torch.bfloat16).reshape(Q * Lq, D).contiguous()
        mask = text_mask.to(dev, torch.float32).contiguous()
        if seg_text_idx is None and Q != V:
            raise ValueError("seg_text_idx is required when the number of texts differs from the number of segments")
        rows = V * T
        x = frames.reshape(rows, D).float()                       # fp32 residual stream
        x_bf = torch.empty((rows, D), dtype=torch.bfloat16, device=dev)
        xp_bf = torch.empty((rows, D), dtype=torch.bfloat16, device=dev)
        eng.layernorm(x, y_pos_bf16=xp_bf, pos=pos, period=T)     # q input of the first layer: x + pos
        qbuf = torch.empty((rows, D), dtype=torch.bfloat16, device=dev)
        att = torch.empty((rows, D), dtype=torch.bfloat16, device=dev)
        hbuf = torch.empty((rows, FFN), dtype=torch.bfloat16, device=dev)
        kv = torch.empty((Q * Lq, 2 * D), dtype=torch.bfloat16, device=dev)
        # ---- 2 x text -> video cross-attention layers
        for i in range(NL):
            pre = f"t2v_encoder.layers.{i}."
            W, B = p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"]
            eng.gemm(xp_bf, W[:D], bias=B[:D], out=qbuf)                                  # Q = (x + pos) Wq^T + bq
            eng.gemm(text2d, W[D:], bias=B[D:], out=kv)                                   # K | V of the text, once per query
            eng.mha96(qbuf, kv[:, :D], kv[:, D:], att, V, NH, T, Lq, kv_seq_idx=seg_text_idx, key_mask=mask)
            eng.gemm(att, p[pre + "self_attn.out_proj.weight"], bias=p[pre + "self_attn.out_proj.bias"], out=x,
                     out_mode=GEMM_ADD_F32)                                               # src2 = x + attn
            eng.layernorm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], y_bf16=x_bf)  # src3 in = norm1(src2)
            eng.gemm(x_bf, p[pre + "linear1.weight"], bias=p[pre + "linear1.bias"], out=hbuf, flags=GEMM_FLAG_RELU)
            eng.gemm(hbuf, p[pre + "linear2.weight"], bias=p[pre + "linear2.bias"], out=x, out_mode=GEMM_ADD_F32)
            eng.layernorm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"], y_f32=x, y_pos_bf16=xp_bf, pos=pos, period=T)
        # ---- prepend the global token; 2 x post-norm self-attention layers over 1 + T tokens
        T1 = T + 1
        rows1 = V * T1
        x1 = torch.empty((V, T1, D), dtype=torch.float32, device=dev)
        x1[:, 0] = p["global_rep_token"].float()
        x1[:, 1:] = x.view(V, T, D)
        x1 = x1.view(rows1, D)
        x1_bf = torch.empty((rows1, D), dtype=torch.bfloat16, device=dev)
        x1p_bf = torch.empty((rows1, D), dtype=torch.bfloat16, device=dev)
        eng.layernorm(x1, y_bf16=x1_bf, y_pos_bf16=x1p_bf, pos=pos_g, period=T1)
        qk = torch.empty((rows1, 2 * D), dtype=torch.bfloat16, device=dev)
        vbuf = torch.empty((rows1, D), dtype=torch.bfloat16, device=dev)
        att1 = torch.empty((rows1, D), dtype=torch.bfloat16, device=dev)
        h1 = torch.empty((rows1, FFN), dtype=torch.bfloat16, device=dev)
        for i in range(NL):
            pre = f"encoder.layers.{i}."
            W, B = p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"]
            eng.gemm(x1p_bf, W[:2 * D], bias=B[:2 * D], out=qk)                           # q = k = x + pos
            eng.gemm(x1_bf, W[2 * D:], bias=B[2 * D:], out=vbuf)                          # v = x
            eng.mha96(qk[:, :D], qk[:, D:], vbuf, att1, V, NH, T1, T1)
            eng.gemm(att1, p[pre + "self_attn.out_proj.weight"], bias=p[pre + "self_attn.out_proj.bias"], out=x1,
                     out_mode=GEMM_ADD_F32)
            eng.layernorm(x1, p[pre + "norm1.weight"], p[pre + "norm1.bias"], y_f32=x1, y_bf16=x1_bf)
            eng.gemm(x1_bf, p[pre + "linear1.weight"], bias=p[pre + "linear1.bias"], out=h1, flags=GEMM_FLAG_RELU)
            eng.gemm(h1, p[pre + "linear2.weight"], bias=p[pre + "linear2.bias"], out=x1, out_mode=GEMM_ADD_F32)
            eng.layernorm(x1, p[pre + "norm2.weight"], p[pre + "norm2.bias"], y_f32=x1, y_bf16=x1_bf, y_pos_bf16=x1p_bf,
                          pos=pos_g, period=T1)
        # ---- CLS row -> Linear(768 -> hidden)
        cls_rows = x1_bf.view(V, T1, D)[:, 0].contiguous()
        return eng.gemm(cls_rows, p["mm_projector.weight"], bias=p["mm_projector.bias"])
