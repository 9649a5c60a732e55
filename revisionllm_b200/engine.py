"""Thin Python owner of one `rvl_handle`: device buffers (torch tensors) + calls into the C ABI.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only - every FLOP of
the hot path runs in the kernels of `csrc/` behind `include/revisionllm_b200.h`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import GEMM_ADD_F32, GEMM_FLAG_RELU, GEMM_FLAG_SWAP, GEMM_OUT_BF16, GEMM_OUT_F32, RvlError


@dataclass
class EngineConfig:
    hidden: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    head_dim: int = 128
    intermediate: int = 11008
    vocab: int = 32000
    adapter_dim: int = 768
    max_pos: int = 4096
    kv_page_size: int = 32
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0

    @classmethod
    def from_synth(cls, s) -> "EngineConfig":
        return cls(hidden=s.hidden, n_layers=s.n_layers, n_heads=s.n_heads, head_dim=s.head_dim,
                   intermediate=s.intermediate, vocab=s.vocab, adapter_dim=s.adapter_dim, max_pos=s.max_pos,
                   rms_eps=s.rms_eps, rope_theta=s.rope_theta)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise RvlError(f"{name}: expected a contiguous CUDA tensor of {dtype}, got {t.dtype} on {t.device}")
    return t


class DecodeChunks:
    """Chunks of greedy decode steps over fixed-address buffers, replayed as CUDA graphs (mixed into `Engine`; the CPU
    stand-in of tests/test_generate_host_cpu.py mixes it in as well, so generate()'s chunk logic is the same code there)."""
    _ws = None
    _kv = None

    _capture_stream = None

    def _init_chunks(self):
        self._dec_bufs: Dict[Tuple[int, int], Dict[str, torch.Tensor]] = {}   # (rows, max_pages) -> fixed-address decode buffers
        self._dec_graphs: Dict[tuple, Tuple[object, int]] = {}                # captured decode chunks (graph, launches per replay)
        self._dec_seen: Dict[tuple, int] = {}

    DECODE_CHUNK = 8          # decode steps per captured graph = steps between two looks at the EOS flags
    GRAPH_AFTER = 3           # capture a chunk shape when it shows up for the third time: capturing + instantiating ~1800 kernel
                              # nodes costs tens of ms (measured: a stage-2 query went from 67 to 144 ms when every call captured),
                              # a replay saves ~0.25 ms per decode step

    def decode_buffers(self, n_rows: int, max_pages: int) -> Dict[str, torch.Tensor]:
        """Fixed-address buffers a captured decode chunk reads and writes (a CUDA graph bakes pointers in): last-row logits,
        sequence lengths, page table, EOS flags, and the chunk's token / entropy rows.  A handful of (rows, pages) shapes is
        kept; the oldest goes first."""
        key = (n_rows, max_pages)
        b = self._dec_bufs.get(key)
        if b is None:
            while len(self._dec_bufs) >= 6:
                old = next(iter(self._dec_bufs))
                del self._dec_bufs[old]
                for gk in [gk for gk in self._dec_graphs if gk[:2] == old]:
                    del self._dec_graphs[gk]
            dev, K = self.device, self.DECODE_CHUNK
            b = dict(logits=torch.empty((n_rows, self.cfg.vocab), dtype=torch.float32, device=dev),
                     seq_lens=torch.empty(n_rows, dtype=torch.int32, device=dev),
                     page_table=torch.empty((n_rows, max_pages), dtype=torch.int32, device=dev),
                     unfinished=torch.empty(n_rows, dtype=torch.int32, device=dev),
                     ring_tok=torch.empty((K, n_rows), dtype=torch.int32, device=dev),
                     ring_ent=torch.empty((K, n_rows), dtype=torch.float32, device=dev))
            self._dec_bufs[key] = b
        return b

    def decode_chunk(self, bufs: Dict[str, torch.Tensor], k: int, eos_id: int, pad_id: int, use_unfinished: bool, max_kv_len: int,
                     graph: bool = True):
        """`k` x (greedy sample of bufs.logits -> ring_tok[s] / ring_ent[s]; decode step on that token -> bufs.logits), the
        generation loop of vtimellm_llama.py:287-369 for k tokens, as ONE CUDA-graph launch once the same chunk shape has been
        seen `GRAPH_AFTER - 1` times before (the first occurrences run eagerly: capture costs more than it saves for a shape
        used once or twice).  Every
        pointer the chunk touches lives in `bufs`, the bound weights, the workspace or the KV pages; the graph is dropped
        when the workspace or the KV pages are re-allocated."""
        n_rows, max_pages = bufs["page_table"].shape
        unf = bufs["unfinished"] if use_unfinished else None

        def body():
            self.decode_n(k, bufs["logits"], bufs["ring_tok"], bufs["ring_ent"], unf, eos_id, pad_id, bufs["seq_lens"], bufs["page_table"],
                          max_kv_len)

        key = (n_rows, max_pages, k, eos_id, pad_id, use_unfinished, max_kv_len, self._ws.data_ptr() if self._ws is not None else 0,
               self._kv.data_ptr() if self._kv is not None else 0)
        hit = self._dec_graphs.get(key)
        if hit is not None:
            hit[0].replay()
            self.launches += hit[1]
            return
        self._dec_seen[key] = self._dec_seen.get(key, 0) + 1
        if not graph or self._dec_seen[key] < self.GRAPH_AFTER or self.device.type != "cuda":
            body()
            return
        for gk in [gk for gk in self._dec_graphs if gk[-2:] != key[-2:]]:      # graphs over buffers that no longer exist
            del self._dec_graphs[gk]
        before = self.launches
        # captured by hand on a side stream: the torch.cuda.graph() context manager also synchronises the device, runs the
        # Python garbage collector and empties the caching allocator, so that the NEXT call pays a cudaMalloc for every tensor
        # it makes (measured: a stage-2 query at 161 instead of 64 ms right after a capture).  Nothing is allocated in `body`.
        g = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream(self.device)
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(device=self.device)
        self._capture_stream.wait_stream(cur)
        with torch.cuda.stream(self._capture_stream):
            g.capture_begin(capture_error_mode="thread_local")
            try:
                body()
            finally:
                g.capture_end()
        cur.wait_stream(self._capture_stream)
        self._dec_graphs[key] = (g, self.launches - before)
        self.launches = before
        g.replay()
        self.launches += self._dec_graphs[key][1]


class Engine(DecodeChunks):
    """One per (process, GPU).  Not thread-safe."""

    def __init__(self, cfg: EngineConfig, device: Optional[int] = None):
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RvlError("revisionllm_b200 needs an sm_100 GPU: torch.cuda.is_available() is False (no CPU fallback)")
        self.cfg = cfg
        self.device_index = torch.cuda.current_device() if device is None else device
        self.device = torch.device("cuda", self.device_index)
        c = _cabi.rvl_config(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate, cfg.vocab,
                             cfg.adapter_dim, cfg.max_pos, cfg.kv_page_size, self.device_index, cfg.rms_eps,
                             cfg.rope_theta)
        h = C.c_void_p()
        _cabi.check(self.lib.rvl_create(C.byref(c), C.byref(h)), None, "rvl_create")
        self.h = h
        self._keep: List[torch.Tensor] = []      # bound weights stay alive as long as the engine
        self._ws: Optional[torch.Tensor] = None
        self._ws_cap = (0, 0)
        self._kv: Optional[torch.Tensor] = None
        self.n_pages = 0
        self.launches = 0                         # kernels enqueued through this engine (bench's gpu_launches)
        self._init_chunks()
        self._clip_ws: Optional[torch.Tensor] = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.rvl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        _cabi.check(rc, self.h, what)

    # ------------------------------------------------------------------ weights
    def bind_state_dict(self, sd: Dict[str, torch.Tensor]):
        """HF-named bf16 state dict -> fused per-layer tensors on this GPU, bound by pointer.
        q/k/v and gate/up are concatenated once here (the 'one-time repack' of SURVEY.md section 8b)."""
        cfg = self.cfg
        dev = self.device

        def g(name):
            return sd[name].to(device=dev, dtype=torch.bfloat16).contiguous()

        layers = (_cabi.rvl_layer_weights * cfg.n_layers)()
        keep: List[torch.Tensor] = []
        # gate/up rows interleaved in blocks of 16 so the GEMM epilogue can apply SwiGLU (rvl_weights.wgu_layout 1)
        I = cfg.intermediate
        interleave = I % 16 == 0 and os.environ.get("RVL_WGU_LAYOUT", "1") != "0"
        for i in range(cfg.n_layers):
            p = f"model.layers.{i}."
            wqkv = torch.cat([g(p + f"self_attn.{n}_proj.weight") for n in ("q", "k", "v")], dim=0).contiguous()
            gate, up = g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")
            if interleave:
                wgu = torch.stack([gate.view(I // 16, 16, -1), up.view(I // 16, 16, -1)], dim=1).reshape(2 * I, -1).contiguous()
            else:
                wgu = torch.cat([gate, up], dim=0).contiguous()
            del gate, up
            wo, wd = g(p + "self_attn.o_proj.weight"), g(p + "mlp.down_proj.weight")
            ln1, ln2 = g(p + "input_layernorm.weight"), g(p + "post_attention_layernorm.weight")
            keep += [wqkv, wgu, wo, wd, ln1, ln2]
            layers[i] = _cabi.rvl_layer_weights(wqkv.data_ptr(), wo.data_ptr(), wgu.data_ptr(), wd.data_ptr(),
                                                ln1.data_ptr(), ln2.data_ptr())
        emb, fn, head = g("model.embed_tokens.weight"), g("model.norm.weight"), g("lm_head.weight")
        keep += [emb, fn, head]
        pw = pb = None
        if "model.mm_projector.weight" in sd:
            pw, pb = g("model.mm_projector.weight"), g("model.mm_projector.bias")
            keep += [pw, pb]
        w = _cabi.rvl_weights(emb.data_ptr(), fn.data_ptr(), head.data_ptr(), _ptr(pw), _ptr(pb), layers, 1 if interleave else 0)
        self._check(self.lib.rvl_bind_weights(self.h, C.byref(w)), "rvl_bind_weights")
        self._keep = keep
        self.embed_tokens, self.lm_head_w, self.proj_w, self.proj_b = emb, head, pw, pb
        self.wgu_interleaved = interleave

    # ------------------------------------------------------------------ buffers
    def ensure_workspace(self, max_tokens: int, max_seqs: int):
        if self._ws is not None and self._ws_cap[0] >= max_tokens and self._ws_cap[1] >= max_seqs:
            return
        max_tokens = max(max_tokens, self._ws_cap[0])
        max_seqs = max(max_seqs, self._ws_cap[1])
        nbytes = self.lib.rvl_workspace_bytes(self.h, max_tokens, max_seqs)
        self._ws = None
        self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._check(self.lib.rvl_set_workspace(self.h, self._ws.data_ptr(), nbytes, max_tokens, max_seqs), "rvl_set_workspace")
        self._ws_cap = (max_tokens, max_seqs)

    def ensure_kv(self, n_pages: int):
        if self._kv is not None and self.n_pages >= n_pages:
            return
        nbytes = self.lib.rvl_kv_bytes(self.h, n_pages)
        self._kv = None
        self._kv = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._check(self.lib.rvl_set_kv(self.h, self._kv.data_ptr(), n_pages), "rvl_set_kv")
        self.n_pages = n_pages

    def kv_bytes(self, n_pages: int) -> int:
        return self.lib.rvl_kv_bytes(self.h, n_pages)

    # ------------------------------------------------------------------ hot path
    def project_splice(self, feats, feat_dst, text_ids, text_dst, hidden_out):
        n_feat = 0 if feats is None else feats.shape[0]
        n_text = 0 if text_ids is None else text_ids.shape[0]
        if n_feat:
            _req(feats, torch.bfloat16, "feats"); _req(feat_dst, torch.int32, "feat_dst")
        if n_text:
            _req(text_ids, torch.int32, "text_ids"); _req(text_dst, torch.int32, "text_dst")
        _req(hidden_out, torch.float32, "hidden_out")
        self._check(self.lib.rvl_project_splice(self.h, _ptr(feats), _ptr(feat_dst), n_feat, _ptr(text_ids),
                                                _ptr(text_dst), n_text, hidden_out.data_ptr(), hidden_out.shape[0],
                                                _stream()), "rvl_project_splice")
        self.launches += (1 if n_feat else 0) + (1 if n_text else 0)

    def gather_windows(self, features_f32, frame_idx):
        """features [T, D] fp32 (device), frame_idx [...] int32 (device) -> bf16 [..., D]."""
        _req(features_f32, torch.float32, "features"); _req(frame_idx, torch.int32, "frame_idx")
        n = frame_idx.numel()
        out = torch.empty(tuple(frame_idx.shape) + (features_f32.shape[1],), dtype=torch.bfloat16, device=features_f32.device)
        self._check(self.lib.rvl_gather_windows(self.h, features_f32.data_ptr(), features_f32.shape[0], features_f32.shape[1],
                                                frame_idx.data_ptr(), n, out.data_ptr(), _stream()), "rvl_gather_windows")
        self.launches += 1
        return out

    def splice_rows(self, vis, vis_dst, text_ids, text_dst, hidden_out):
        n_vis = 0 if vis is None else vis.shape[0]
        n_text = 0 if text_ids is None else text_ids.shape[0]
        if n_vis:
            _req(vis, torch.bfloat16, "vis"); _req(vis_dst, torch.int32, "vis_dst")
        _req(hidden_out, torch.float32, "hidden_out")
        self._check(self.lib.rvl_splice_rows(self.h, _ptr(vis), _ptr(vis_dst), n_vis, _ptr(text_ids), _ptr(text_dst),
                                             n_text, hidden_out.data_ptr(), hidden_out.shape[0], _stream()),
                    "rvl_splice_rows")
        self.launches += (1 if n_vis else 0) + (1 if n_text else 0)

    def prefill(self, hidden, cu_seqlens, n_seq, max_seqlen, page_table, logits_out, all_logits=False, seq_pos0=None,
                seq_ctx_row=None):
        _req(hidden, torch.float32, "hidden"); _req(cu_seqlens, torch.int32, "cu_seqlens")
        _req(page_table, torch.int32, "page_table"); _req(logits_out, torch.float32, "logits_out")
        T = hidden.shape[0]
        self.ensure_workspace(T, n_seq)
        self._check(self.lib.rvl_prefill(self.h, hidden.data_ptr(), cu_seqlens.data_ptr(), n_seq, T, max_seqlen,
                                         page_table.data_ptr(), page_table.shape[1], logits_out.data_ptr(),
                                         1 if all_logits else 0, _ptr(seq_pos0), _ptr(seq_ctx_row), _stream()), "rvl_prefill")
        self.launches += 1 + (8 if self.wgu_interleaved else 9) * self.cfg.n_layers + 2 + (0 if all_logits else 1)   # + last-row gather

    def decode_step(self, token_ids, seq_lens, page_table, logits_out, max_kv_len: int = 0):
        _req(token_ids, torch.int32, "token_ids"); _req(seq_lens, torch.int32, "seq_lens")
        _req(page_table, torch.int32, "page_table"); _req(logits_out, torch.float32, "logits_out")
        n = token_ids.shape[0]
        self.ensure_workspace(n, n)
        self._check(self.lib.rvl_decode_step(self.h, token_ids.data_ptr(), seq_lens.data_ptr(), n, page_table.data_ptr(),
                                             page_table.shape[1], max_kv_len, logits_out.data_ptr(), _stream()), "rvl_decode_step")
        self.launches += 1 + (7 if self.wgu_interleaved else 8) * self.cfg.n_layers + 3

    def decode_n(self, n_steps, logits, token_ring, entropy_ring, unfinished, eos_id, pad_id, seq_lens, page_table, max_kv_len: int = 0):
        """`n_steps` x (greedy sample of `logits` -> token_ring[s] / entropy_ring[s]; decode step on that token -> `logits`) in one
        C call (rvl_decode_n)."""
        _req(logits, torch.float32, "logits"); _req(token_ring, torch.int32, "token_ring"); _req(seq_lens, torch.int32, "seq_lens")
        _req(page_table, torch.int32, "page_table")
        n = logits.shape[0]
        if token_ring.shape[0] < n_steps or token_ring.shape[1] != n:
            raise RvlError("decode_n: token_ring must be [>= n_steps, n_seq]")
        self.ensure_workspace(n, n)
        self._check(self.lib.rvl_decode_n(self.h, n_steps, logits.data_ptr(), token_ring.data_ptr(), _ptr(entropy_ring), _ptr(unfinished),
                                          eos_id, pad_id, seq_lens.data_ptr(), n, page_table.data_ptr(), page_table.shape[1], max_kv_len,
                                          _stream()), "rvl_decode_n")
        self.launches += n_steps * (2 + (7 if self.wgu_interleaved else 8) * self.cfg.n_layers + 3)

    def sample_greedy(self, logits, next_tokens, entropy=None, unfinished=None, eos_id=2, pad_id=2):
        _req(logits, torch.float32, "logits"); _req(next_tokens, torch.int32, "next_tokens")
        n, v = logits.shape
        self._check(self.lib.rvl_sample_greedy(self.h, logits.data_ptr(), n, v, _ptr(unfinished), eos_id, pad_id,
                                               next_tokens.data_ptr(), _ptr(entropy), _stream()), "rvl_sample_greedy")
        self.launches += 1

    def sample_multinomial(self, logits, next_tokens, temperature, seed, step, entropy=None, unfinished=None, eos_id=2, pad_id=2,
                           philox_out=None):
        _req(logits, torch.float32, "logits"); _req(next_tokens, torch.int32, "next_tokens")
        n, v = logits.shape
        self._check(self.lib.rvl_sample_multinomial(self.h, logits.data_ptr(), n, v, float(temperature), int(seed) & (2 ** 64 - 1),
                                                    int(step), _ptr(unfinished), eos_id, pad_id, next_tokens.data_ptr(),
                                                    _ptr(entropy), _ptr(philox_out), _stream()), "rvl_sample_multinomial")
        self.launches += 1

    def cosine_topk(self, frames, seg_offsets, cls, k=3, norm_axis=1, max_seg_rows=None, want_idx=True, seg_ends=None):
        """seg_ends None: proposal i = rows [seg_offsets[i], seg_offsets[i + 1]); else rows [seg_offsets[i], seg_ends[i])."""
        _req(frames, torch.bfloat16, "frames"); _req(seg_offsets, torch.int32, "seg_offsets"); _req(cls, torch.bfloat16, "cls")
        n_seg = seg_offsets.shape[0] - 1
        if seg_ends is not None:
            _req(seg_ends, torch.int32, "seg_ends")
            n_seg = seg_ends.shape[0]
        if max_seg_rows is None:
            max_seg_rows = int(frames.shape[0])
        scores = torch.empty(n_seg, dtype=torch.float32, device=frames.device)
        idx = torch.empty((n_seg, k), dtype=torch.int32, device=frames.device) if want_idx else None
        self._check(self.lib.rvl_cosine_topk(self.h, frames.data_ptr(), seg_offsets.data_ptr(), _ptr(seg_ends), n_seg, frames.shape[1],
                                             cls.data_ptr(), k, norm_axis, min(max_seg_rows, 8192), scores.data_ptr(),
                                             _ptr(idx), _stream()), "rvl_cosine_topk")
        self.launches += 1
        return scores, idx

    def select_topk(self, scores, k):
        _req(scores, torch.float32, "scores")
        idx = torch.empty(k, dtype=torch.int32, device=scores.device)
        self._check(self.lib.rvl_select_topk(self.h, scores.data_ptr(), scores.shape[0], k, idx.data_ptr(), _stream()),
                    "rvl_select_topk")
        self.launches += 1
        return idx

    def merge_rank(self, cos, ent, keep, cover1=None, cover_all=None, mode=0, normalize=True, minmax=True):
        """-> (scores fp64 [n], order int32 [n], n_out int32 [1]) on the device (see rvl_merge_rank)."""
        n = keep.shape[0]
        for t, name in ((cos, "cos"), (ent, "ent")):
            if t is not None:
                _req(t, torch.float32, name)
        for t, name in ((keep, "keep"), (cover1, "cover1"), (cover_all, "cover_all")):
            if t is not None:
                _req(t, torch.int32, name)
        scores = torch.empty(n, dtype=torch.float64, device=keep.device)
        order = torch.full((n,), -1, dtype=torch.int32, device=keep.device)
        n_out = torch.zeros(1, dtype=torch.int32, device=keep.device)
        self._check(self.lib.rvl_merge_rank(self.h, _ptr(cos), _ptr(ent), keep.data_ptr(), _ptr(cover1), _ptr(cover_all), n, mode,
                                            1 if normalize else 0, 1 if minmax else 0, scores.data_ptr(), order.data_ptr(),
                                            n_out.data_ptr(), _stream()), "rvl_merge_rank")
        self.launches += 1
        return scores, order, n_out

    # ------------------------------------------------------------------ measurement
    def profile(self, on: bool, capacity: int = 16384):
        self._check(self.lib.rvl_profile_enable(self.h, 1 if on else 0, capacity), "rvl_profile_enable")

    def profile_read(self, category: int):
        ms, fl, by, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        self._check(self.lib.rvl_profile_read(self.h, category, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)), "rvl_profile_read")
        return dict(ms=ms.value, flops=fl.value, bytes=by.value, launches=n.value)

    # ------------------------------------------------------------------ single kernels
    def gemm(self, A, W, bias=None, out=None, out_mode=GEMM_OUT_BF16, flags=0, rowmap=None, split_k=1, ldc=None):
        _req(A, torch.bfloat16, "A"); _req(W, torch.bfloat16, "W")
        M, K = A.shape
        N = W.shape[0]
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16 if out_mode == GEMM_OUT_BF16 else torch.float32, device=A.device)
        ldc = out.shape[1] if ldc is None else ldc
        self._check(self.lib.rvl_gemm_bf16(self.h, A.data_ptr(), W.data_ptr(), _ptr(bias), out.data_ptr(), M, N, K, ldc,
                                           out_mode, flags, _ptr(rowmap), split_k, _stream()), "rvl_gemm_bf16")
        self.launches += 1
        return out

    def rmsnorm(self, x, w, eps=None, rows=None):
        _req(x, torch.float32, "x"); _req(w, torch.bfloat16, "w")
        n = x.shape[0] if rows is None else rows.shape[0]
        y = torch.empty((n, x.shape[1]), dtype=torch.bfloat16, device=x.device)
        self._check(self.lib.rvl_rmsnorm(self.h, x.data_ptr(), w.data_ptr(), y.data_ptr(), n, x.shape[1],
                                         self.cfg.rms_eps if eps is None else eps, _ptr(rows), _stream()), "rvl_rmsnorm")
        self.launches += 1
        return y

    def rope_kv(self, qkv, page_table, layer, positions=None, tok_seq=None, cu_seqlens=None):
        _req(qkv, torch.bfloat16, "qkv")
        self._check(self.lib.rvl_rope_kv(self.h, qkv.data_ptr(), qkv.shape[0], _ptr(positions), _ptr(tok_seq),
                                         _ptr(cu_seqlens), page_table.data_ptr(), page_table.shape[1], layer, _stream()),
                    "rvl_rope_kv")
        self.launches += 1

    def swiglu(self, gu):
        _req(gu, torch.bfloat16, "gu")
        inter = gu.shape[1] // 2
        act = torch.empty((gu.shape[0], inter), dtype=torch.bfloat16, device=gu.device)
        self._check(self.lib.rvl_swiglu(self.h, gu.data_ptr(), act.data_ptr(), gu.shape[0], inter, _stream()), "rvl_swiglu")
        self.launches += 1
        return act

    def attn_prefill(self, qkv, cu_seqlens, n_seq, max_seqlen):
        _req(qkv, torch.bfloat16, "qkv")
        out = torch.empty((qkv.shape[0], self.cfg.hidden), dtype=torch.bfloat16, device=qkv.device)
        self._check(self.lib.rvl_attn_prefill(self.h, qkv.data_ptr(), out.data_ptr(), cu_seqlens.data_ptr(), n_seq,
                                              max_seqlen, qkv.shape[0], _stream()), "rvl_attn_prefill")
        self.launches += 1
        return out

    def attn_decode(self, qkv, seq_lens, page_table, layer, fused_rope=False, max_kv_len: int = 0):
        _req(qkv, torch.bfloat16, "qkv")
        out = torch.empty((qkv.shape[0], self.cfg.hidden), dtype=torch.bfloat16, device=qkv.device)
        self._check(self.lib.rvl_attn_decode(self.h, qkv.data_ptr(), out.data_ptr(), seq_lens.data_ptr(), qkv.shape[0],
                                             page_table.data_ptr(), page_table.shape[1], layer, 1 if fused_rope else 0, max_kv_len, _stream()),
                    "rvl_attn_decode")
        self.launches += 1
        return out

    def clip_encoder(self, weights, frames, text, text_mask, seg_text_idx, out):
        """rvl_clip_encoder: the whole stage-2 adapter in one C call.  `weights`: a filled _cabi.rvl_clip_weights."""
        _req(frames, torch.bfloat16, "frames"); _req(text, torch.bfloat16, "text"); _req(text_mask, torch.float32, "text_mask")
        V, T, _ = frames.shape
        Q, Lq, _ = text.shape
        nbytes = self.lib.rvl_clip_encoder_workspace_bytes(V, T, Q, Lq)
        if self._clip_ws is None or self._clip_ws.numel() < nbytes:
            self._clip_ws = None
            self._clip_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._check(self.lib.rvl_clip_encoder(self.h, C.byref(weights), frames.data_ptr(), text.data_ptr(), text_mask.data_ptr(),
                                              _ptr(seg_text_idx), V, T, Q, Lq, self._clip_ws.data_ptr(), self._clip_ws.numel(),
                                              out.data_ptr(), _stream()), "rvl_clip_encoder")
        self.launches += 3 + 2 * 8 + 1 + 2 * 8 + 2 + 1
        return out

    def layernorm(self, x, w=None, b=None, y_f32=None, y_bf16=None, pos=None, y_pos_bf16=None, period=0, eps=1e-5):
        _req(x, torch.float32, "x")
        rows, dim = x.shape
        self._check(self.lib.rvl_layernorm(self.h, x.data_ptr(), _ptr(w), _ptr(b), _ptr(y_f32), _ptr(y_bf16), _ptr(pos),
                                           _ptr(y_pos_bf16), rows, dim, period, eps, _stream()), "rvl_layernorm")
        self.launches += 1

    def mha96(self, q, k, v, out, n_seq, n_heads, Tq, Tk, kv_seq_idx=None, key_mask=None):
        """q/k/v/out: 2-D bf16 views whose rows are `stride(0)` apart (column slices of fused projections are fine)."""
        for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
            if t.dtype != torch.bfloat16 or not t.is_cuda or t.stride(1) != 1:
                raise RvlError(f"mha96: {n} must be a CUDA bf16 matrix with unit column stride")
        self._check(self.lib.rvl_mha96(self.h, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                       out.data_ptr(), out.stride(0), n_seq, n_heads, Tq, Tk, _ptr(kv_seq_idx),
                                       k.shape[0] // Tk, _ptr(key_mask), _stream()), "rvl_mha96")
        self.launches += 1

    def kv_view(self, layer: int):
        """(k, v) views [n_pages, n_heads, page, 128] bf16 of one layer's cache (tests only)."""
        c = self.cfg
        per = self.n_pages * c.n_heads * c.kv_page_size * c.head_dim
        flat = self._kv.view(torch.bfloat16)
        k = flat[(2 * layer) * per:(2 * layer + 1) * per].view(self.n_pages, c.n_heads, c.kv_page_size, c.head_dim)
        v = flat[(2 * layer + 1) * per:(2 * layer + 2) * per].view(self.n_pages, c.n_heads, c.kv_page_size, c.head_dim)
        return k, v


# ---------------------------------------------------------------------- host-side splice planning
def plan_splice(input_ids: np.ndarray, n_visual: Sequence[int], attention_mask: Optional[np.ndarray] = None,
                max_length: Optional[int] = None, image_token: int = -200):
    """Index math of `prepare_inputs_labels_for_multimodal`
    (/root/reference/revisionllm/model/vtimellm_arch.py:149-244) without touching embeddings:
    for each row drop masked ids, replace every placeholder by that row's next visual block, truncate to
    `max_length`, and pack all rows back to back.

    input_ids [B, Ltxt] int64; n_visual[i] = rows of the i-th visual block (blocks are consumed in
    order, one per placeholder; a row without placeholder consumes one block and uses none of it).
    Returns dict(cu_seqlens [B+1], text_ids, text_dst, vis_src, vis_dst, lengths) as int32 numpy arrays;
    vis_src indexes rows of the concatenated visual blocks."""
    B = input_ids.shape[0]
    vis_start = np.concatenate([[0], np.cumsum(np.asarray(n_visual, dtype=np.int64))])
    valid = np.ones(input_ids.shape, dtype=bool) if attention_mask is None else attention_mask.astype(bool)
    is_img = (input_ids == image_token) & valid
    if (is_img.sum(axis=1) == 1).all() and len(n_visual) == B:
        # common case (one placeholder per row, block i belongs to row i): pure numpy, no per-token Python loop
        nv = (vis_start[1:] - vis_start[:-1]).astype(np.int64)                      # [B]
        tok_len = np.where(is_img, nv[:, None], valid.astype(np.int64))             # rows each id expands to
        start = np.cumsum(tok_len, axis=1) - tok_len                                 # position inside the row
        limit = max_length if max_length is not None else np.iinfo(np.int64).max
        row_len = np.minimum(tok_len.sum(axis=1), limit)
        cu64 = np.concatenate([[0], np.cumsum(row_len)])
        txt = valid & ~is_img & (start < limit)
        rows_t, cols_t = np.nonzero(txt)
        text_ids = input_ids[rows_t, cols_t]
        text_dst = cu64[rows_t] + start[rows_t, cols_t]
        img_col = np.argmax(is_img, axis=1)
        img_start = start[np.arange(B), img_col]
        n_keep = np.clip(np.minimum(nv, limit - img_start), 0, None)                 # visual rows surviving truncation
        rep = np.repeat(np.arange(B), n_keep)
        within = np.arange(int(n_keep.sum())) - np.repeat(np.cumsum(n_keep) - n_keep, n_keep)
        vis_src = vis_start[rep] + within
        vis_dst = cu64[rep] + img_start[rep] + within
        i32v = lambda a: np.asarray(a, dtype=np.int32)
        return dict(cu_seqlens=i32v(cu64), text_ids=i32v(text_ids), text_dst=i32v(text_dst), vis_src=i32v(vis_src),
                    vis_dst=i32v(vis_dst), lengths=i32v(row_len))
    text_ids: List[int] = []
    text_dst: List[int] = []
    vis_src: List[int] = []
    vis_dst: List[int] = []
    cu = [0]
    blk = 0
    for b in range(B):
        ids = input_ids[b]
        if attention_mask is not None:
            ids = ids[attention_mask[b].astype(bool)]
        base = cu[-1]
        pos = 0
        limit = max_length if max_length is not None else 1 << 60
        n_img = int((ids == image_token).sum())
        if n_img == 0:
            blk += 1                       # vtimellm_arch.py:168-176
        for t in ids.tolist():
            if t == image_token:
                n = int(vis_start[blk + 1] - vis_start[blk])
                for r in range(n):
                    if pos < limit:
                        vis_src.append(int(vis_start[blk]) + r)
                        vis_dst.append(base + pos)
                        pos += 1
                blk += 1
            else:
                if pos < limit:
                    text_ids.append(int(t))
                    text_dst.append(base + pos)
                    pos += 1
        cu.append(base + pos)
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    lengths = np.diff(np.asarray(cu, dtype=np.int64)).astype(np.int32)
    return dict(cu_seqlens=i32(cu), text_ids=i32(text_ids), text_dst=i32(text_dst), vis_src=i32(vis_src),
                vis_dst=i32(vis_dst), lengths=lengths)
