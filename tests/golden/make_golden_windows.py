"""Golden window tables produced by EXECUTING the reference's own source lines (run where /root/reference exists):

    python tests/golden/make_golden_windows.py

The window builders are inline blocks of two evaluation scripts, not functions:
  * stage 1: /root/reference/revisionllm/eval/eval_nlq_negative.py:224-241 (`ctx_l = len(features)` ... `plus_baseline`);
  * stage 2: /root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:262-294 (windows, times, stage-1 positives ->
    `grounding_windows`, even padding with non-grounding windows).
The blocks are cut out by their first / last statements, dedented and exec'd with synthetic `args`, `features` (row t holds
the value t, so the gathered rows ARE the indices), `grounding_dict` and `batch`.  Output: tests/golden/windows.json."""
import json
import math
import os
import textwrap
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/revisionllm/eval/"


def cut(path, first, last):
    lines = open(path).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == first)
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip() == last)
    return textwrap.dedent("\n".join(lines[i0:i1 + 1])), (i0 + 1, i1 + 1)


def main():
    s1_src, s1_lines = cut(REF + "eval_nlq_negative.py", "ctx_l = len(features)", "clip_feats.append(clip_feat)")
    # the first `clip_feats.append` closes the window loop; the plus_baseline block follows it
    lines = open(REF + "eval_nlq_negative.py").read().split("\n")
    end = s1_lines[1]
    extra = next(i for i in range(end, len(lines)) if lines[i].strip() == "clip_feats.append(clip_feat)")
    s1_src = textwrap.dedent("\n".join(lines[s1_lines[0] - 1: extra + 1]))
    s1_lines = (s1_lines[0], extra + 1)
    s2_src, s2_lines = cut(REF + "eval_nlq_retrieval_e2e2.py", "ctx_l = len(features)", "grounding_windows = list(range(len(clip_feats)))")
    out = {"source": {"stage1": f"eval_nlq_negative.py:{s1_lines[0]}-{s1_lines[1]}", "stage2": f"eval_nlq_retrieval_e2e2.py:{s2_lines[0]}-{s2_lines[1]}"},
           "stage1": [], "stage2": []}
    for (T, window, fps, nf, baseline, plus) in [(18000, 100, 5, 250, False, False), (18000, 100, 5, 100, False, True), (1234, 60, 5, 100, False, False),
                                                 (601, 100, 5, 64, False, False), (999, 125, 2, 100, False, True), (18000, 100, 5, 250, True, False)]:
        ns = {"np": np, "math": math, "features": np.arange(T, dtype=np.int64)[:, None],
              "args": SimpleNamespace(debug_window=window, feature_fps=fps, num_frames=nf, baseline=baseline, plus_baseline=plus)}
        exec(s1_src, ns)
        out["stage1"].append(dict(ctx_l=T, debug_window=window, feature_fps=fps, num_frames=nf, baseline=baseline, plus_baseline=plus,
                                  windows=[w[:, 0].tolist() for w in ns["clip_feats"]]))
    rng = np.random.default_rng(0)
    for (T, window, fps, nf, stride, batch, p_pos) in [(18000, 100, 5, 250, 5, 100, 0.2), (18000, 100, 5, 250, 5, 33, 0.1), (9000, 100, 5, 100, 2, 20, 0.3),
                                                      (18000, 100, 5, 250, 5, 100, 0.9), (3100, 60, 5, 50, 5, 12, 0.25), (18000, 100, 5, 250, 5, 100, None)]:
        clip_length = window * fps
        n_stage1 = math.ceil(T / (clip_length // 2)) - 1
        answers = None if p_pos is None else ["From 3 to 9" if rng.random() < p_pos else "Not Present" for _ in range(n_stage1)]
        ns = {"np": np, "math": math, "features": np.arange(T, dtype=np.int64)[:, None], "id": "q0", "batch": batch,
              "grounding_dict": {} if answers is None else {"q0": {"answer": answers}},
              "args": SimpleNamespace(debug_window=window, feature_fps=fps, num_frames=nf, stride=stride)}
        exec(s2_src, ns)
        out["stage2"].append(dict(ctx_l=T, debug_window=window, feature_fps=fps, num_frames=nf, stride=stride, batch=batch, stage1_answers=answers,
                                  times=[[int(a), int(b)] for a, b in ns["times"]], grounding_windows=[int(i) for i in ns["grounding_windows"]],
                                  selected_first_frames=[int(w[0, 0]) for w in ns["clip_feats"]],
                                  n_windows=len(ns["windowidx"])))
    # random small cases: every stride, batches below / at / above the number of positive windows
    rng2 = np.random.default_rng(1)
    while len(out["stage2"]) < 46:
        window, fps = int(rng2.integers(5, 40)), int(rng2.integers(1, 6))
        clip_length = window * fps
        stride = int(rng2.integers(2, 7))
        if clip_length // stride == 0 or clip_length // 2 == 0:
            continue
        T = int(rng2.integers(clip_length + 2, 40 * clip_length))
        nf = int(rng2.integers(2, 12))
        batch = int(rng2.integers(1, 60))
        n_stage1 = math.ceil(T / (clip_length // 2)) - 1
        p_pos = float(rng2.uniform(0, 1))
        answers = ["From 3 to 9" if rng2.random() < p_pos else "Not Present" for _ in range(n_stage1)]
        ns = {"np": np, "math": math, "features": np.arange(T, dtype=np.int64)[:, None], "id": "q0", "batch": batch,
              "grounding_dict": {"q0": {"answer": answers}},
              "args": SimpleNamespace(debug_window=window, feature_fps=fps, num_frames=nf, stride=stride)}
        try:
            exec(s2_src, ns)
        except (ValueError, IndexError):          # slice step 0 / window id outside the list: the reference skips the query (bare except, :338)
            continue
        out["stage2"].append(dict(ctx_l=T, debug_window=window, feature_fps=fps, num_frames=nf, stride=stride, batch=batch, stage1_answers=answers,
                                  times=[[int(a), int(b)] for a, b in ns["times"]], grounding_windows=[int(i) for i in ns["grounding_windows"]],
                                  selected_first_frames=[int(w[0, 0]) for w in ns["clip_feats"]],
                                  n_windows=len(ns["windowidx"])))
    json.dump(out, open(os.path.join(HERE, "windows.json"), "w"))
    print(out["source"], [len(c["windows"]) for c in out["stage1"]], [(c["n_windows"], len(c["grounding_windows"])) for c in out["stage2"]])


if __name__ == "__main__":
    main()
