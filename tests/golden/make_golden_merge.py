"""Golden merged scores produced by EXECUTING the reference's own in-script merge (run where /root/reference exists):

    python tests/golden/make_golden_merge.py

/root/reference/revisionllm/eval/eval_nlq_negative.py:321-335 (`if args.normalize:` ... `scores = score_cos`: max-normalisation
of the cosine and entropy scores, then `cos - entropy`, `cos / entropy`, `-entropy` or the cosine score alone) is cut out of
the script's main loop, dedented and exec'd on seeded score lists.  Output: tests/golden/merge_inline.json."""
import json
import os
import textwrap
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/revisionllm/eval/eval_nlq_negative.py"


def main():
    lines = open(SRC).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == "if args.normalize:")
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip() == "scores = score_cos")
    src = textwrap.dedent("\n".join(lines[i0:i1 + 1]))
    rng = np.random.default_rng(3)
    cases = []
    for n in (1, 2, 7, 57):
        cos = [float(v) for v in rng.uniform(0.05, 0.9, size=n)]
        ent = [float(v) for v in rng.uniform(0.2, 3.0, size=n)]
        for normalize in (True, False):
            for score, merge in (("entropy", "add"), ("entropy", "multiply"), ("entropy", "none"), ("cosine_sim", "add")):
                ns = dict(args=SimpleNamespace(normalize=normalize, score=score, score_merge=merge), score_cos=list(cos), scores_entropy=list(ent))
                exec(src, ns)
                cases.append(dict(score_cos=cos, scores_entropy=ent, normalize=normalize, score=score, score_merge=merge,
                                  scores=[float(v) for v in ns["scores"]]))
    json.dump(dict(source=f"eval_nlq_negative.py:{i0 + 1}-{i1 + 1}", cases=cases), open(os.path.join(HERE, "merge_inline.json"), "w"))
    print(f"lines {i0 + 1}-{i1 + 1}", len(cases), "cases")


if __name__ == "__main__":
    main()
