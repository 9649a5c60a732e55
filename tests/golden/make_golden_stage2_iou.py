"""Golden answers of the stage-2 script's own `iou` (run where /root/reference exists):

    python tests/golden/make_golden_stage2_iou.py

`iou` is lifted out of /root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:109-139 with `ast` (the script opens
LMDB environments and parses arguments at import time) and called on synthetic answers, chunk starts (negative ones
included: fewer windows than one chunk), permutations, zoom levels and grounding-window lists.
Output: tests/golden/stage2_iou.json."""
import ast
import json
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py"


def main():
    fn = next(n for n in ast.parse(open(SRC).read()).body if isinstance(n, ast.FunctionDef) and n.name == "iou")
    ns = {"re": re}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    ref_iou = ns["iou"]
    rng = np.random.default_rng(7)
    cases = []
    for n_windows, batch in [(100, 100), (33, 33), (155, 100), (12, 100), (20, 100), (7, 4)]:
        starts, indexes, zooms = [], [], []
        for zoom in (4, 2, 1):
            b = batch // zoom
            for i in range(int(np.ceil(n_windows / b))):
                start = i * b
                end = min(start + b, n_windows)
                if end - start < b:
                    start = end - b
                n = len(range(n_windows)[start:end])
                starts.append(int(start)); indexes.append([int(v) for v in rng.permutation(n)]); zooms.append(zoom)
        gw = sorted(int(v) for v in rng.choice(179, size=n_windows, replace=False))
        answers = []
        for i in range(len(starts)):
            r = rng.random()
            answers.append("Not Present" if r < 0.15 else (f"Video {int(rng.integers(0, 130))}." if r < 0.8 else f"From {int(rng.integers(0, 400))} to 3"))
        a, b_ = sorted(float(v) for v in rng.uniform(0, 179, size=2))
        gt = [round(a, 2), round(b_, 2)] if len(cases) % 2 else [round(b_, 2), round(max(0.0, a - 40.0), 2)]   # also given as (end, start)
        frames, hit = ref_iou(answers, gt, 250, batch, starts, indexes, True, zooms, gw)
        cases.append(dict(outputs=answers, gt=gt, num_frames_video=batch, starts=starts, indexes=indexes, hierarchy_zooms=zooms,
                          grounding_windows=gw, clip_frames={str(k): [int(v[0]), int(v[1])] for k, v in frames.items()}, hit=hit))
    json.dump(dict(source="eval_nlq_retrieval_e2e2.py:109-139", cases=cases), open(os.path.join(HERE, "stage2_iou.json"), "w"))
    print([(len(c["outputs"]), len(c["clip_frames"]), c["hit"]) for c in cases])


if __name__ == "__main__":
    main()
