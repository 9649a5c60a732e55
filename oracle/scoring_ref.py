"""Scoring, windowing and selection rules restated (TEST INFRASTRUCTURE).

Follows:
  * /root/reference/revisionllm/eval/similarity.py:71-94             (_topk_pooling)
  * /root/reference/revisionllm/eval/eval_nlq_negative.py:309-336    (stage-1 cosine score + merge)
  * /root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:380-386 (stage-2 cosine score)
  * /root/reference/revisionllm/uncertainty/funs_get_feature_X.py:120-146 (get_entropy_statistics)
  * /root/reference/revisionllm/eval/eval_nlq_negative.py:224-235    (stage-1 windows, 50% overlap)
  * /root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:262-294 (stage-2 windows + selection)
  * /root/reference/revisionllm/eval/eval_nlq_negative.py:79-112     (answer parsing)
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------- cosine top-k
def topk_lowest_index(x: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k largest values, descending, ties -> lowest index.
    (torch.topk leaves tie order unspecified; the build pins it)"""
    order = np.lexsort((np.arange(x.shape[0]), -x.astype(np.float64)))
    return order[:k]


def topk_pooling(text_embeds: torch.Tensor, video_embeds: torch.Tensor, k: int) -> torch.Tensor:
    """similarity.py:71-94. text [Nt, D], video [Nv, F, D] -> [Nv, Nt, D]:
    sims = video @ text.T, top-k frames per (video, text), SUM of those frames."""
    Nt, D = text_embeds.shape
    sims = (video_embeds.float() @ text_embeds.float().t()).numpy()       # [Nv, F, Nt]
    out = torch.zeros(video_embeds.shape[0], Nt, D)
    for v in range(video_embeds.shape[0]):
        for t in range(Nt):
            idx = topk_lowest_index(sims[v, :, t], k)
            out[v, t] = video_embeds[v, torch.from_numpy(idx.copy())].float().sum(0)
    return out


def cosine_topk_score(frames: torch.Tensor, cls: torch.Tensor, k: int = 3, norm_axis: int = 1
                      ) -> Tuple[float, np.ndarray, np.ndarray]:
    """Score of one proposal = dot(sum of its top-k normalised frames, cls).

    norm_axis=1: per-frame L2 norm (stage 2, eval_nlq_retrieval_e2e2.py:382).
    norm_axis=0: the stage-1 driver's quirk - `.norm(dim=0)` over the FRAME axis
    (eval_nlq_negative.py:311).  `cls` is used un-normalised (:209).
    Returns (score, top-k frame indices, all sims).  fp32 throughout."""
    f = frames.float()
    f = f / f.norm(dim=norm_axis, keepdim=True)
    sims = (f @ cls.float()).numpy().astype(np.float32)
    kk = min(k, f.shape[0])
    idx = topk_lowest_index(sims, kk)
    pooled = f[torch.from_numpy(idx.copy())].sum(0)
    return float(torch.dot(pooled, cls.float())), idx, sims


def select_topk_segments(scores: np.ndarray, k: int) -> np.ndarray:
    """BASELINE.json north_star: stage-2 segments chosen by per-segment cosine
    score top-k; bit-exact given identical fp32 scores (ties -> lowest index)."""
    return topk_lowest_index(np.asarray(scores, dtype=np.float32), min(k, len(scores)))


# --------------------------------------------------------------------------- entropy
def get_entropy_statistics(logits: torch.Tensor, q_begin: int = 0, q_end: Optional[int] = None) -> torch.Tensor:
    """funs_get_feature_X.py:120-146. logits [B, T, V] -> [B, 4] = (max, min,
    mean, std) over steps of H_t = -sum p*log(p + 1e-10); std is unbiased and 0
    when there is a single step."""
    if q_end is None:
        q_end = logits.shape[1]
    probs = torch.softmax(logits[:, q_begin:q_end, :].float(), dim=2)
    ent = -torch.sum(probs * torch.log(probs + 1e-10), dim=2)
    if q_end == q_begin + 1:
        std = torch.zeros(ent.shape[0])
    else:
        std = ent.std(dim=1)
    return torch.stack([ent.max(dim=1).values, ent.min(dim=1).values, ent.mean(dim=1), std], dim=1)


def step_entropy(logits: torch.Tensor) -> torch.Tensor:
    """Per-step entropies [B] for one [B, V] score tensor (same formula)."""
    p = torch.softmax(logits.float(), dim=-1)
    return -torch.sum(p * torch.log(p + 1e-10), dim=-1)


# --------------------------------------------------------------------------- windows
def stage1_windows(ctx_l: int, clip_length: int, num_frames: int) -> np.ndarray:
    """eval_nlq_negative.py:224-235: 50%-overlap windows, np.linspace int32
    sample indices [W, num_frames]."""
    num_window = math.ceil(ctx_l / (clip_length // 2)) - 1
    out = []
    for i in range(num_window):
        start = max(i * clip_length // 2, 0)
        end = min(i * clip_length // 2 + clip_length, ctx_l - 1)
        out.append(np.linspace(start, end, num_frames, dtype=np.int32))
    return np.array(out, dtype=np.int32).reshape(len(out), num_frames)


def stage2_windows(ctx_l: int, clip_length: int, num_frames: int, stride: int = 5) -> Tuple[np.ndarray, List[Tuple[int, int]]]:
    """eval_nlq_retrieval_e2e2.py:262-277: stride clip_length//stride, the
    last windows are clamped to full length (start = end - clip_length)."""
    num_window = math.ceil(ctx_l / (clip_length // stride)) - 1
    out, times = [], []
    for i in range(num_window):
        start = max(i * clip_length // stride, 0)
        end = min(i * clip_length // stride + clip_length, ctx_l - 1)
        if end - start < clip_length:
            start = end - clip_length
        times.append((start, end))
        out.append(np.linspace(start, end, num_frames, dtype=np.int32))
    return np.array(out, dtype=np.int32).reshape(len(out), num_frames), times


def nonoverlap_segments(ctx_l: int, num_frames: int) -> np.ndarray:
    """BASELINE.json config 2: non-overlapping `num_frames`-frame segments
    (18000 features -> 180 x 100)."""
    n = ctx_l // num_frames
    return np.arange(n * num_frames, dtype=np.int32).reshape(n, num_frames)


def stage2_select_windows(stage1_answers: Sequence[str], n_stage2_windows: int, batch: int, stride: int = 5) -> List[int]:
    """eval_nlq_retrieval_e2e2.py:278-294: stage-1 windows whose answer is not
    'Not Present', mapped from the stride-2 grid to the stride-`stride` grid,
    de-duplicated with `list(set(...))` exactly as the reference does (:284 - when no padding follows, the list keeps
    CPython's set iteration order, which is a function of the inserted values only), padded with evenly spaced other
    windows up to `batch` and sorted in that case (:285-290)."""
    gw: List[int] = []
    for i, a in enumerate(stage1_answers):
        if a != "Not Present":
            lo = math.floor((i - 1) * (stride / 2))
            hi = math.ceil((i - 1) * (stride / 2) + (stride / 2))
            gw.extend(range(lo, hi))
    gw = list(set(gw))
    if batch > len(gw):
        non = [i for i in range(n_stage2_windows) if i not in gw]
        if len(non) > 0:
            step = int(len(non) / (batch - len(gw)))
            non = non[::step][: batch - len(gw)] if step > 0 else non[: batch - len(gw)]
        gw = sorted(gw + non)
    return gw


# --------------------------------------------------------------------------- answers
_SPAN = re.compile(r"(\d+) (to|and) (\d+)")


def parse_span(answer: str) -> Optional[Tuple[int, int]]:
    """eval_nlq_negative.py:87-92: first '(\\d+) (to|and) (\\d+)' match, ordered."""
    m = _SPAN.search(answer)
    if not m:
        return None
    a, b = int(m.group(1)), int(m.group(3))
    return (min(a, b), max(a, b))


def merge_scores(score_cos: Sequence[float], score_ent: Sequence[float], mode: str = "add", normalize: bool = True) -> List[float]:
    """eval_nlq_negative.py:321-336: divide each list by its max, then
    cos - entropy ('add') or cos / entropy ('multiply')."""
    c, e = list(score_cos), list(score_ent)
    if normalize:
        if c:
            m = max(c)
            c = [x / m for x in c]
        if e:
            m = max(e)
            e = [x / m for x in e]
    if mode == "add":
        return [a - b for a, b in zip(c, e)]
    if mode == "multiply":
        return [a / b for a, b in zip(c, e)]
    return [-x for x in e]
