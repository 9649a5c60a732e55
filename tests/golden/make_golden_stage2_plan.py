"""Golden call plans of the stage-2 zoom recursion, produced by EXECUTING the reference's own loop (run where /root/reference
exists):

    python tests/golden/make_golden_stage2_plan.py

/root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:337-353 - `for hierarchy_zoom in [4,2,1]` ... `inference(...)`,
`hierarchy_zooms.append(...)` - is cut out of the script's main loop, dedented and exec'd with `features[w]` = w (so the
visual input of every call shows which windows it holds), a recording stub for `inference` and `torch.manual_seed(seed)` in
front (the reference draws its permutations from the global generator, :348).  Output: tests/golden/stage2_plan.json -
per case the chunk starts, permutations and the window ids of every generate() call, in call order."""
import json
import math
import os
import textwrap
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py"


def main():
    lines = open(SRC).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == "for hierarchy_zoom in [4,2,1]:")
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip() == "hierarchy_zooms.append(hierarchy_zoom)")
    src = textwrap.dedent("\n".join(lines[i0:i1 + 1]))
    cases = []
    for (n_windows, batch, seed) in [(100, 100, 0), (33, 33, 1), (100, 33, 2), (155, 100, 3), (7, 4, 4), (12, 100, 5)]:
        calls = []

        def inference(model, feat, query_feats, prompt, tokenizer, return_list=False):
            calls.append([int(v) for v in feat[0, :, 0, 0]])
            return ["0"], {}
        ns = dict(math=math, torch=torch, args=SimpleNamespace(batch=batch, q_feat_dir=None), inference=inference, model=None, tokenizer=None,
                  features=torch.arange(n_windows, dtype=torch.float32)[:, None, None].expand(n_windows, 2, 3).contiguous(),
                  query_feats_temp=None, query="{}", sentence="s", answers=[], starts=[], indexes=[], hierarchy_zooms=[])
        torch.manual_seed(seed)
        exec(src, ns)
        cases.append(dict(n_windows=n_windows, batch=batch, seed=seed, starts=[int(s) for s in ns["starts"]],
                          perms=[[int(v) for v in p] for p in ns["indexes"]], zooms=[int(z) for z in ns["hierarchy_zooms"]], call_windows=calls))
    json.dump({"source": f"eval_nlq_retrieval_e2e2.py:{i0 + 1}-{i1 + 1}", "cases": cases}, open(os.path.join(HERE, "stage2_plan.json"), "w"))
    print(f"lines {i0 + 1}-{i1 + 1}", [(c["n_windows"], c["batch"], len(c["call_windows"])) for c in cases])


if __name__ == "__main__":
    main()
