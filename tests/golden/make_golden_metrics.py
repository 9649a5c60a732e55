"""Generates tests/golden/merge_metrics.json from the REFERENCE itself (run in the build container, where
/root/reference exists):
  * writes synthetic stage-1 / stage-2 prediction files in the reference's JSONL format
    (eval_nlq_negative.py:115-125 `write_log`), runs /root/reference/revisionllm/eval/metric_retrieval_forward.py on them
    as a subprocess and records the metrics it writes to result_retrieval.txt;
  * exec-s the reference's own `iou` function (extracted from eval_nlq_negative.py with `ast`, because the module itself
    imports lmdb / vtimellm which are absent) on seeded answers and records its outputs.
Nothing of the reference's source is copied into the repo: only inputs and outputs are stored.
"""
import ast
import json
import os
import random
import re
import subprocess
import sys
import tempfile

REF = "/root/reference/revisionllm/eval"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref_iou():
    src = open(os.path.join(REF, "eval_nlq_negative.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "iou")
    ns = {"re": re}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_iou", "exec"), ns)
    return ns["iou"]


def answers(rng, n, clip=250):
    out = []
    for _ in range(n):
        r = rng.random()
        if r < 0.45:
            out.append("Not Present")
        elif r < 0.55:
            out.append(f"From {clip - 1} to {clip - 1}.")
        elif r < 0.65:
            a = rng.randrange(clip)
            out.append(f"From {a} to {a}.")
        else:
            a = rng.randrange(clip - 1)
            out.append(f"From {a} to {rng.randrange(a, clip)}.")
    return out


def main():
    rng = random.Random(7)
    ref_iou = load_ref_iou()
    iou_cases = []
    for case in range(6):
        n = rng.choice([4, 17, 57])
        ans = answers(rng, n)
        scores = [round(rng.uniform(0.5, 4.0), 4) for _ in range(n)] if case != 2 else []
        gt = sorted([rng.random(), rng.random()])
        nfv = int(rng.uniform(2000, 9000))
        plus = case == 4
        cf, ious, fs = ref_iou(ans, tuple(gt), 250, nfv, scores, plus)
        iou_cases.append(dict(answers=ans, gt=gt, num_frames_clip=250, num_frames_video=nfv, scores=scores, plus_baseline=plus,
                              clip_frames={str(k): list(v) for k, v in cf.items()}, ious=ious, kept_scores=fs))
    # ---- merge + ranking metrics through the reference script
    queries = []
    with tempfile.TemporaryDirectory() as tmp:
        for d in ("g", "r1", "r2"):
            os.makedirs(os.path.join(tmp, d))
        for q in range(24):
            n = rng.choice([20, 57, 143])
            ans = answers(rng, n)
            nfv = n * 125 + 125
            gt = sorted([rng.random(), rng.random()])
            spans = [(i, re.search(r"(\d+) to (\d+)", a)) for i, a in enumerate(ans)]
            spans = [(i, int(m.group(1)), int(m.group(2))) for i, m in spans if m and not (int(m.group(1)) == 249 == int(m.group(2)))]
            if spans and q % 4 != 1:                            # ground truth near one of the proposals, so recalls are not all zero
                i, a, b = rng.choice(spans)
                gt = [(i * 125 + a + rng.randrange(-5, 6)) / nfv, (i * 125 + b + rng.randrange(0, 30)) / nfv]
                gt = [max(0.0, gt[0]), max(gt[0] + 1e-3, gt[1])]
            ent = [round(rng.uniform(0.2, 3.0), 5) for _ in range(n)]
            cf, ious, fs = ref_iou(ans, tuple(gt), 250, nfv, ent, False)
            cos = [round(rng.uniform(0.1, 0.9), 5) for _ in ious]
            if q % 5 == 0 and len(cos) > 1:
                cos[1] = cos[0]
                fs[1] = fs[0]                                   # exact tie: the ranking must keep index order
            top_c, top_e = (max(cos) if cos else 1), (max(fs) if fs else 1)
            scores = [c / top_c - e / top_e for c, e in zip(cos, fs)]              # --normalize, --score_merge add (:321-331)
            gl = dict(video_id=f"m{q % 3}", task="grounding", query_id=q, answer=ans, info=dict(iou=ious, scores=scores))

            def retrieval(k):
                fr = {}
                for j in range(k):
                    lo = rng.randrange(0, max(1, int(2.5 * n) - 10))
                    fr[str(j)] = [lo, lo + rng.randrange(1, 40)]
                return dict(video_id=gl["video_id"], task="grounding", query_id=q, answer=["x"] * k,
                            info=dict(frames=fr, mean_entropy=[round(rng.uniform(0.3, 2.0), 4) for _ in range(max(k, 2))], score_cos=[]))
            rl, rl2 = retrieval(rng.choice([1, 3, 6])), retrieval(rng.choice([1, 2]))
            if q == 3:
                rl["info"]["frames"] = {"0": [10 ** 6, 10 ** 6 + 5]}          # covers nothing -> the log passes through unfiltered
            queries.append(dict(gl=gl, rl=rl, rl2=rl2, cos=cos, ent=fs))
            for d, log in (("g", gl), ("r1", rl), ("r2", rl2)):
                with open(os.path.join(tmp, d, "predictions_negative_0.txt"), "a") as f:
                    f.write(json.dumps(log) + "\n")
        r = subprocess.run([sys.executable, os.path.join(REF, "metric_retrieval_forward.py"), "--grounding_path", os.path.join(tmp, "g"),
                            "--retrieval_path", os.path.join(tmp, "r1"), "--retrieval_path2", os.path.join(tmp, "r2"),
                            "--distributed_grounding", "1", "--distributed_retrieval", "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        metrics = json.load(open(os.path.join(tmp, "g", "result_retrieval.txt")))
        selected_ratio = float([l for l in r.stdout.splitlines() if re.fullmatch(r"[0-9.]+", l.strip())][0])
    json.dump(dict(iou_cases=iou_cases, queries=queries, metrics=metrics, selected_ratio=selected_ratio),
              open(os.path.join(HERE, "merge_metrics.json"), "w"))
    print("wrote merge_metrics.json:", {k: round(v, 3) for k, v in list(metrics.items())[:6]}, "selected", selected_ratio)


if __name__ == "__main__":
    main()
