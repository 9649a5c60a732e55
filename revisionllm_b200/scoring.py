"""Scoring utilities of the eval drivers: windows, cosine top-k, entropy statistics, selection, merge.

Device work goes through the C ABI (`rvl_cosine_topk`, `rvl_select_topk`, `rvl_sample_greedy`);
the host functions mirror the drivers' integer bookkeeping:
  * /root/reference/revisionllm/eval/eval_nlq_negative.py:224-235   stage-1 windows (50 % overlap)
  * /root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:262-294 stage-2 windows + selection
  * /root/reference/revisionllm/eval/similarity.py:71-94            `_topk_pooling`
  * /root/reference/revisionllm/uncertainty/funs_get_feature_X.py:120-146 `get_entropy_statistics`
  * /root/reference/revisionllm/eval/eval_nlq_negative.py:79-112,321-336 answer parsing, score merge
"""
from __future__ import annotations

import math
import re
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import Engine


# ------------------------------------------------------------------------------------ windows (host, integers)
def _linspace_i32(start: int, end: int, n: int) -> np.ndarray:
    return np.linspace(start, end, n, dtype=np.int32)


def stage1_windows(ctx_l: int, clip_length: int, num_frames: int) -> np.ndarray:
    """[W, num_frames] int32 frame indices; W = ceil(ctx_l / (clip_length // 2)) - 1."""
    half = clip_length // 2
    n_win = math.ceil(ctx_l / half) - 1
    rows = [_linspace_i32(max(i * clip_length // 2, 0), min(i * clip_length // 2 + clip_length, ctx_l - 1), num_frames)
            for i in range(n_win)]
    return np.stack(rows).astype(np.int32) if rows else np.zeros((0, num_frames), np.int32)


def stage2_windows(ctx_l: int, clip_length: int, num_frames: int, stride: int = 5):
    """Stride clip_length // stride; trailing windows are shifted left to keep the full length."""
    step = clip_length // stride
    n_win = math.ceil(ctx_l / step) - 1
    rows, times = [], []
    for i in range(n_win):
        s = max(i * clip_length // stride, 0)
        e = min(i * clip_length // stride + clip_length, ctx_l - 1)
        if e - s < clip_length:
            s = e - clip_length
        times.append((s, e))
        rows.append(_linspace_i32(s, e, num_frames))
    return (np.stack(rows).astype(np.int32) if rows else np.zeros((0, num_frames), np.int32)), times


def nonoverlap_segments(ctx_l: int, num_frames: int) -> np.ndarray:
    """BASELINE.json config 2: consecutive `num_frames`-frame segments (18000 feats -> 180 x 100)."""
    n = ctx_l // num_frames
    return np.arange(n * num_frames, dtype=np.int32).reshape(n, num_frames)


def stage2_select_windows(stage1_answers: Sequence[str], n_stage2_windows: int, batch: int, stride: int = 5) -> List[int]:
    """Positive stage-1 windows mapped to the stride-`stride` grid, padded with evenly spaced others
    (eval_nlq_retrieval_e2e2.py:278-294).  Order: sorted when padding was needed (:289); otherwise the order in which a
    Python set of the mapped ids iterates, exactly as the reference's `list(set(...))` leaves it (:284) - ids below zero
    (a positive stage-1 window 0) index from the end of the window list, as they do there."""
    mapped: List[int] = []
    for i, ans in enumerate(stage1_answers):
        if ans == "Not Present":
            continue
        lo = math.floor((i - 1) * (stride / 2))
        hi = math.ceil((i - 1) * (stride / 2) + (stride / 2))
        mapped.extend(range(lo, hi))
    picked = list(set(mapped))
    missing = batch - len(picked)
    if missing > 0:
        rest = [i for i in range(n_stage2_windows) if i not in picked]
        if rest:
            step = int(len(rest) / missing)
            rest = rest[::step][:missing] if step > 0 else rest[:missing]
        picked = sorted(picked + rest)
    return picked


_SPAN_RE = re.compile(r"(\d+) (to|and) (\d+)")
_NUM_RE = re.compile(r"(\d+)")


def parse_span(answer: str) -> Optional[Tuple[int, int]]:
    m = _SPAN_RE.search(answer)
    if m is None:
        return None
    a, b = int(m.group(1)), int(m.group(3))
    return (a, b) if a <= b else (b, a)


def parse_first_int(answer: str) -> Optional[int]:
    m = _NUM_RE.search(answer)
    return None if m is None else int(m.group(1))


def merge_scores(score_cos: Sequence[float], score_ent: Sequence[float], mode: str = "add", normalize: bool = True) -> List[float]:
    cos, ent = list(score_cos), list(score_ent)
    if normalize:
        if cos:
            top = max(cos)
            cos = [c / top for c in cos]
        if ent:
            top = max(ent)
            ent = [e / top for e in ent]
    if mode == "add":
        return [c - e for c, e in zip(cos, ent)]
    if mode == "multiply":
        return [c / e for c, e in zip(cos, ent)]
    return [-e for e in ent]


# ------------------------------------------------------------------------------------ device
NORM_PER_FRAME, NORM_ACROSS_FRAMES, NORM_NONE = 1, 0, 2


def cosine_topk_scores(engine: Engine, frames: torch.Tensor, seg_offsets: torch.Tensor, cls: torch.Tensor, k: int = 3,
                       norm_axis: int = NORM_PER_FRAME, max_seg_rows: Optional[int] = None):
    """frames [n_rows, D] bf16 (all proposals back to back), seg_offsets [n_seg+1] int32, cls [D] bf16.
    Returns (scores [n_seg] fp32, top-k frame indices [n_seg, k] int32, -1 padded)."""
    return engine.cosine_topk(frames, seg_offsets, cls, k=k, norm_axis=norm_axis, max_seg_rows=max_seg_rows)


def _topk_pooling(engine: Engine, text_embeds: torch.Tensor, video_embeds: torch.Tensor, k: int) -> torch.Tensor:
    """Drop-in for similarity.py:71-94: text [Nt, D], video [Nv, F, D] -> pooled [Nv, Nt, D] (sum of the
    top-k frames by raw dot product).  Index selection runs in the CUDA kernel; the gather+sum of k rows is
    plumbing."""
    Nv, F, D = video_embeds.shape
    vb = video_embeds.to(engine.device, torch.bfloat16).reshape(Nv * F, D).contiguous()
    offs = torch.arange(0, (Nv + 1) * F, F, dtype=torch.int32, device=engine.device)
    out = torch.empty((Nv, text_embeds.shape[0], D), dtype=torch.float32, device=engine.device)
    for t in range(text_embeds.shape[0]):
        _, idx = engine.cosine_topk(vb, offs, text_embeds[t].to(engine.device, torch.bfloat16).contiguous(), k=k,
                                    norm_axis=NORM_NONE, max_seg_rows=F)
        rows = (idx.long() + (offs[:-1].long())[:, None]).reshape(-1)
        out[:, t] = vb.float().index_select(0, rows).view(Nv, k, D).sum(1)
    return out


def select_topk_segments(engine: Engine, scores: torch.Tensor, k: int) -> torch.Tensor:
    return engine.select_topk(scores.to(engine.device, torch.float32).contiguous(), min(k, scores.shape[0]))


def get_entropy_statistics(engine: Engine, logits: torch.Tensor, q_begin: int = 0, q_end: Optional[int] = None) -> torch.Tensor:
    """Drop-in for funs_get_feature_X.py:120-146: logits [B, T, V] -> [B, 4] = (max, min, mean, std)."""
    B, T, V = logits.shape
    q_end = T if q_end is None else q_end
    ent = torch.empty((q_end - q_begin, B), dtype=torch.float32, device=engine.device)
    tok = torch.empty(B, dtype=torch.int32, device=engine.device)
    for i, t in enumerate(range(q_begin, q_end)):
        engine.sample_greedy(logits[:, t].to(engine.device, torch.float32).contiguous(), tok, ent[i])
    return entropy_stats_from_steps(ent.t())


def entropy_stats_from_steps(ent: torch.Tensor) -> torch.Tensor:
    """ent [B, T'] per-step entropies (already produced on the device during decode) -> [B, 4] = (max, min, mean, std) as
    in funs_get_feature_X.py:136-145.  NaN marks steps a row did not run (retired after EOS): they are left out."""
    valid = ~torch.isnan(ent)
    cnt = valid.sum(dim=1).clamp(min=1).to(ent.dtype)
    zero = torch.zeros((), dtype=ent.dtype, device=ent.device)
    mx = torch.where(valid, ent, torch.full_like(ent, float("-inf"))).max(dim=1).values
    mn = torch.where(valid, ent, torch.full_like(ent, float("inf"))).min(dim=1).values
    mean = torch.where(valid, ent, zero).sum(dim=1) / cnt
    dev2 = torch.where(valid, (ent - mean[:, None]) ** 2, zero).sum(dim=1)
    std = torch.sqrt(dev2 / (cnt - 1).clamp(min=1)) * (cnt > 1).to(ent.dtype)          # unbiased, 0 for a single step
    return torch.stack([mx, mn, mean, std], dim=1)
