"""TEST INFRASTRUCTURE ONLY (CPU oracle) - never imported by the product path.

Plain-Python restatement of the reference's answer -> proposal -> score -> rank rules:
  * `iou`                      /root/reference/revisionllm/eval/eval_nlq_negative.py:79-112
  * `merge_scores`             ... eval_nlq_negative.py:317-336 (lives in oracle/scoring_ref.py)
  * `stage2_iou`               /root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:109-139 (pinned by
                               tests/golden/stage2_iou.json: the reference's own function, lifted with `ast`)
  * `merge_with_retrieval`     /root/reference/revisionllm/eval/metric_retrieval_forward.py:104-177 (the `--single` branch,
                               buffer = 0; the script's main block, restated as a function)
  * `grounding_metrics_stream` /root/reference/revisionllm/eval/metric_retrieval_forward.py:35-56
Pinned by tests/golden/merge_metrics.json, which tests/golden/make_golden_metrics.py produced by running the
reference script itself on synthetic prediction files (and by exec-ing the reference's own `iou`).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence, Tuple

_SPAN = re.compile(r"(\d+) (to|and) (\d+)")


def iou(outputs: Sequence[str], gt: Tuple[float, float], num_frames_clip: int, num_frames_video: int,
        scores: Sequence[float], plus_baseline: bool = False):
    """-> (clip_frames {window: (from, to)}, ious [n kept], scores of the kept windows)."""
    spans: List[Tuple[int, int]] = []
    kept_scores: List[float] = []
    clip_frames: Dict[int, Tuple[int, int]] = {}
    for idx, text in enumerate(outputs):
        i = 0 if (plus_baseline and idx == len(outputs) - 1) else idx      # :85-86
        m = _SPAN.search(text)
        if not m:
            continue
        a, b = float(m.group(1)), float(m.group(3))
        if a == num_frames_clip - 1 and b == num_frames_clip - 1:          # 'From 249 to 249.' sentinel (:91-92)
            continue
        if a == b:                                                         # :93-95
            a = max(0, a - 1)
            b = min(num_frames_video, b + 1)
        clip_frames[i] = (int(a), int(b))
        spans.append((int(i * num_frames_clip // 2 + a), int(i * num_frames_clip // 2 + b)))   # windows overlap by half
        if len(scores) > 0:
            kept_scores.append(scores[i])
    s, e = gt
    ious = []
    for f0, t0 in spans:
        f, t = f0 / num_frames_video, t0 / num_frames_video
        inter = max(0, min(t, e) - max(f, s))
        union = max(t, e) - min(f, s)
        ious.append(round(inter / union, 2))
    return clip_frames, ious, kept_scores


def stage2_iou(outputs: Sequence[str], gt: Sequence[float], num_frames_video: int, starts: Sequence[int],
               indexes: Sequence[Sequence[int]], hierarchy_zooms: Sequence[int], grounding_windows: Sequence[int]):
    """/root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:109-139, statement by statement."""
    frames = []
    clip_frames = {}
    for i, output in enumerate(outputs):
        m = re.search(r"(\d+)", output)
        if m:
            n = int(m.group(1))
            n = n // hierarchy_zooms[i]                                   # :117
            if n < len(indexes[i]):                                       # :118-119
                n = int(indexes[i][n])
            n = starts[i] + n                                             # :120
            n = max(0, n)
            n = min(len(grounding_windows) - 1, n)
            n = grounding_windows[n]                                      # :123
            to = n
            n = max(0, n - 1)
            to = min(num_frames_video, to + 1)
            clip_frames[i] = (int(n), int(to))
            frames.append((n, to))
    s, e = min(gt), max(gt)
    hits = [max(0, min(t, e) - max(f, s)) for f, t in frames]
    return clip_frames, [1] if sum(hits) > 0 else [0]


def merge_with_retrieval(gl: dict, rl: dict, rl2: Optional[dict], buffer: int = 0) -> dict:
    """One grounding log (stage 1) filtered by the windows the stage-2 logs kept; returns the (mutated copy of the) log."""
    gl = {**gl, "info": dict(gl["info"])}
    n = len(gl["answer"])
    gl_idx = [i for i, a in enumerate(gl["answer"]) if a != "Not Present" and a != "From 249 to 249."]

    def cover(log):
        fr: List[int] = []
        for lo, hi in list(log["info"]["frames"].values()):
            fr.extend(range(max(0, int(.4 * lo) - buffer), min(int(.4 * hi) + buffer, n - 1)))
        return fr

    frames = cover(rl) if len(rl["answer"]) > 0 else []                    # :119-123 (outer loop only repeats the same extend)
    present_idx1 = [i for i in gl_idx if i in frames]
    if rl2 is not None and "frames" in rl2["info"]:
        frames = frames + cover(rl2)
    frames = list(set(frames))
    present_idx = [i for i in gl_idx if i in frames]
    if len(present_idx1) > 0:
        answer = [gl["answer"][i] for i in present_idx]
        ious = [gl["info"]["iou"][gl_idx.index(i)] for i in present_idx]
        sc = gl["info"]["scores"]
        if len(sc) > 0:                                                     # min-max over ALL stage-1 proposals (:146-152)
            lo, hi = min(sc), max(sc)
            if lo != hi:
                sc = [(x - lo) / (hi - lo) for x in sc]
                gl["info"]["scores"] = sc
        scores = [sc[gl_idx.index(i)] for i in present_idx]
        if any(a != "Not Present" for a in answer):
            gl["answer"], gl["info"]["iou"], gl["info"]["scores"] = answer, ious, scores
    return gl


def grounding_metrics_stream(all_logs: Sequence[dict]) -> Dict[str, float]:
    import numpy as np
    ranked = []
    for log in all_logs:
        try:
            sc = log["info"]["scores"]
            order = sorted(range(len(sc)), key=lambda k: sc[k], reverse=True)     # stable: ties keep index order
            ranked.append(np.array([log["info"]["iou"][i] for i in order]))
        except Exception:
            ranked.append(np.array([log["info"]["iou"]]))
    n = len(ranked)
    if n == 0:
        return {}
    out: Dict[str, float] = {"mIoU": sum(u[0] for u in ranked if len(u) >= 1) / n * 100}
    for m in [0.1, 0.3, 0.5, 0.7, 0.9]:
        for r in [1, 5, 10, 50]:
            out[f"R{r}@{m}"] = 0.0
        for u in ranked:
            hit = u > m
            for r in [1, 5, 10, 50]:
                out[f"R{r}@{m}"] += hit[:r].any() / n * 100
    return out
