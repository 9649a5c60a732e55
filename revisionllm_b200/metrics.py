"""Answer strings -> proposals -> merged scores -> ranked list -> recall metrics.

Host-side mirror of the reference's evaluation tail, written for this code base:
  * `iou`                      /root/reference/revisionllm/eval/eval_nlq_negative.py:79-112
  * `stage2_frames`            /root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:109-139 (the stage-2 script's `iou`)
  * `cover_mask`               /root/reference/revisionllm/eval/metric_retrieval_forward.py:119-135 (windows kept by stage 2)
  * `rank_query`               eval_nlq_negative.py:317-336 + metric_retrieval_forward.py:137-160, numeric part on the GPU
                               (`rvl_merge_rank`)
  * `grounding_metrics_stream` metric_retrieval_forward.py:35-56
String parsing stays on the host (as in the reference); normalisation, merge, filter and ranking run in one kernel.
"""
from __future__ import annotations

import re
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

_SPAN = re.compile(r"(\d+) (to|and) (\d+)")
MODES = {"add": 0, "multiply": 1, "entropy": 2, "cosine": 3}


def parse_proposals(outputs: Sequence[str], num_frames_clip: int, num_frames_video: int, plus_baseline: bool = False):
    """-> (window index, clip-local (from, to), global (from, to)) for every answer that names a span."""
    rows = []
    last = len(outputs) - 1
    for idx, text in enumerate(outputs):
        m = _SPAN.search(text)
        if m is None:
            continue
        a, b = float(m.group(1)), float(m.group(3))
        if a == num_frames_clip - 1 and b == num_frames_clip - 1:
            continue                                    # the model's "nothing here" sentinel
        if a == b:
            a, b = max(0, a - 1), min(num_frames_video, b + 1)
        w = 0 if (plus_baseline and idx == last) else idx
        off = w * num_frames_clip // 2                  # consecutive windows overlap by half a clip
        rows.append((w, (int(a), int(b)), (int(off + a), int(off + b))))
    return rows


def iou(outputs: Sequence[str], gt: Tuple[float, float], num_frames_clip: int, num_frames_video: int, scores: Sequence[float],
        plus_baseline: bool = False):
    """Same return value as the reference's `iou`: ({window: (from, to)}, [iou per proposal], [score per proposal])."""
    rows = parse_proposals(outputs, num_frames_clip, num_frames_video, plus_baseline)
    clip_frames = {w: local for w, local, _ in rows}
    s, e = gt
    ious = []
    for _, _, (f0, t0) in rows:
        f, t = f0 / num_frames_video, t0 / num_frames_video
        inter = max(0, min(t, e) - max(f, s))
        ious.append(round(inter / (max(t, e) - min(f, s)), 2))
    kept = [scores[w] for w, _, _ in rows] if len(scores) > 0 else []
    return clip_frames, ious, kept


_INT = re.compile(r"(\d+)")


def stage2_frames(outputs: Sequence[str], gt: Sequence[float], num_frames_video: int, starts: Sequence[int],
                  indexes: Sequence[Sequence[int]], hierarchy_zooms: Sequence[int], grounding_windows: Sequence[int]):
    """The stage-2 script's `iou` (/root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:109-139): the first integer
    of answer i, divided by the call's zoom, indexes the call's permuted chunk; `starts[i] +` that, clamped, indexes
    `grounding_windows`; the window w becomes the frame range (max(0, w - 1), min(num_frames_video, w + 1)) of the stage-2
    log (`clip_frames`, what `cover_mask` consumes), and the query counts as a hit when any range overlaps [min(gt), max(gt)].
    Same return value as the reference: ({call: (from, to)}, [1] or [0]).  `starts` / `indexes` / `hierarchy_zooms` are the
    per-call lists of the zoom loop (`sweep.stage2_pass` returns them as `start` / `perm` / `zoom`)."""
    clip_frames: Dict[int, Tuple[int, int]] = {}
    s, e = min(gt), max(gt)
    overlap = 0
    for i, text in enumerate(outputs):
        m = _INT.search(text)
        if m is None:
            continue
        j = int(m.group(1)) // hierarchy_zooms[i]
        if j < len(indexes[i]):
            j = int(indexes[i][j])
        j = min(len(grounding_windows) - 1, max(0, starts[i] + j))
        w = int(grounding_windows[j])
        lo, hi = max(0, w - 1), min(num_frames_video, w + 1)
        clip_frames[i] = (lo, hi)
        overlap += max(0, min(hi, e) - max(lo, s))
    return clip_frames, [1] if overlap > 0 else [0]


def cover_mask(n_windows: int, stage2_frames: Dict, buffer: int = 0) -> np.ndarray:
    """Stage-1 windows covered by the windows a stage-2 log kept: stage-2 frame f maps to stage-1 window int(0.4 f)."""
    mask = np.zeros(n_windows, dtype=np.int32)
    for lo, hi in stage2_frames.values():
        a, b = max(0, int(.4 * lo) - buffer), min(int(.4 * hi) + buffer, n_windows - 1)
        if b > a:
            mask[a:b] = 1
    return mask


def rank_query(engine, answers: Sequence[str], cos: Sequence[float], ent: Sequence[float], ious: Sequence[float],
               stage2_frames: Optional[Dict] = None, stage2_frames_b: Optional[Dict] = None, mode: str = "add",
               normalize: bool = True, minmax: bool = True) -> Dict[str, list]:
    """One query: `answers` per stage-1 window; `cos` / `ent` / `ious` per PROPOSAL (the windows whose answer is a span, in
    window order - what `iou` returns).  Returns the ranked proposals: window ids, ious and scores, best first."""
    n = len(answers)
    present = [i for i, a in enumerate(answers) if a != "Not Present" and a != "From 249 to 249."]
    if len(present) != len(ious):
        raise ValueError("one iou per answered window expected")
    dev = engine.device
    keep = np.zeros(n, dtype=np.int32)
    keep[present] = 1
    c = np.zeros(n, dtype=np.float32); c[present] = np.asarray(cos, dtype=np.float32)
    e = np.zeros(n, dtype=np.float32); e[present] = np.asarray(ent, dtype=np.float32)
    cov1 = cov_all = None
    if stage2_frames is not None:
        m1 = cover_mask(n, stage2_frames)
        m_all = np.maximum(m1, cover_mask(n, stage2_frames_b)) if stage2_frames_b is not None else m1
        cov1, cov_all = torch.from_numpy(m1).to(dev), torch.from_numpy(m_all).to(dev)
    scores, order, n_out = engine.merge_rank(torch.from_numpy(c).to(dev), torch.from_numpy(e).to(dev), torch.from_numpy(keep).to(dev),
                                             cov1, cov_all, MODES[mode], normalize, minmax)
    k = int(n_out.item())
    order = order[:k].cpu().numpy()
    sc = scores.cpu().numpy()
    pos = {w: j for j, w in enumerate(present)}
    return dict(windows=order.tolist(), scores=[float(sc[w]) for w in order], ious=[ious[pos[int(w)]] for w in order])


def grounding_metrics_stream(ranked_ious: Sequence[Sequence[float]]) -> Dict[str, float]:
    """mIoU of the top proposal and R{1,5,10,50}@{0.1..0.9} over queries whose proposals are already ranked best first."""
    n = len(ranked_ious)
    if n == 0:
        return {}
    out = {"mIoU": sum(u[0] for u in ranked_ious if len(u) >= 1) / n * 100}
    for m in (0.1, 0.3, 0.5, 0.7, 0.9):
        for r in (1, 5, 10, 50):
            out[f"R{r}@{m}"] = sum(bool((np.asarray(u[:r]) > m).any()) / n * 100 for u in ranked_ious)
    return out
