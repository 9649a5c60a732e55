"""Golden behaviour of the reference's own `inference()` (run where /root/reference exists):

    python tests/golden/make_golden_inference.py

`inference` is lifted out of /root/reference/revisionllm/inference.py:28-75 with `ast` (the module imports clip, easydict,
decord, ... at import time) and exec'd with the reference's own `conv_templates`, `SeparatorStyle`, `tokenizer_image_token`
and `IMAGE_TOKEN_INDEX` (through oracle/ref_shim.py), a recording fake model and the stub tokenizer.  For each case the
golden stores what the function passed to `model.generate` (ids, batch, keyword arguments) and what it returned for the
scripted new tokens (decoded, stop-string and white-space handling, list vs. single string)."""
import ast
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from revisionllm_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/revisionllm/inference.py"


class FakeModel:
    def __init__(self, new_tokens):
        self.new_tokens = torch.tensor(new_tokens)
        self.seen = None

    def generate(self, input_ids, **kw):
        self.seen = dict(input_ids=input_ids.cpu().tolist(), images_shape=list(kw["images"].shape),
                         kwargs={k: v for k, v in kw.items() if k not in ("images", "query_feats", "visual_memory", "prefix_memory")},
                         has_query_feats=kw.get("query_feats") is not None, has_visual_memory=kw.get("visual_memory") is not None)
        return {"sequences": torch.cat([input_ids.cpu(), self.new_tokens], dim=1), "scores": ()}


class _Stop:
    def __init__(self, keywords, tokenizer, input_ids):
        pass


def main():
    ref = ref_shim.load()
    fn = next(n for n in ast.parse(open(SRC).read()).body if isinstance(n, ast.FunctionDef) and n.name == "inference")
    ns = dict(torch=torch, conv_templates=ref.conversation.conv_templates, SeparatorStyle=ref.conversation.SeparatorStyle,
              tokenizer_image_token=ref.mm_utils.tokenizer_image_token, IMAGE_TOKEN_INDEX=ref.constants.IMAGE_TOKEN_INDEX,
              KeywordsStoppingCriteria=_Stop, print=lambda *a, **k: None)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    ref_inference = ns["inference"]
    tok = syn.StubTokenizer(32000)
    f = tok.fixed
    word = tok._word
    stop = word("</s>")                                   # the v1 template's sep2 as the stub tokenizer encodes it ...
    tok.inv[stop] = "</s>"                                # ... and decodes it (a literal "</s>" that skip_special_tokens leaves in the text)
    cases = []
    scripts = [
        ("<video>\nDuring which frames can we see a dog?", 1, [[f["From"], f["1"], f["2"], f["to"], f["3"], f["4"], f["."], 2, 0, 0]], False, False),
        ("<video>\nDuring which frames can we see a dog?", 3, [[f["Not"], f["Present"], 2, 0], [f["From"], f["7"], f["to"], f["9"]], [2, 0, 0, 0]], False, False),
        ("<video>\nDuring which video can we see a cat?", 1, [[f["3"], f["."], 2]], True, False),
        ("<video>\nWhere is the cat?", 2, [[f["From"], f["0"], f["to"], f["5"]], [f["Not"], f["Present"], 2, 2]], False, True),
        ("<video>\nsay the stop string", 1, [[f["Not"], f["Present"], stop, 2]], False, False),
    ]
    for query, B, new, as_list, memory in scripts:
        model = FakeModel(new)
        image = torch.zeros(B, 4, 8)
        vm = torch.zeros(B, 2, 8) if memory else None
        pm = torch.zeros(B, 3, dtype=torch.long) if memory else None
        out, mo = ref_inference(model, image, None, query, tok, visual_memory=vm, prefix_memory=pm, return_list=as_list)
        cases.append(dict(query=query, batch=B, new_tokens=new, return_list=as_list, memory=memory, outputs=out,
                          generate=model.seen, sequences=mo["sequences"].tolist()))
    json.dump(dict(source="revisionllm/inference.py:28-75", cases=cases), open(os.path.join(HERE, "inference_surface.json"), "w"))
    print([c["outputs"] for c in cases])
    print(cases[0]["generate"]["kwargs"])


if __name__ == "__main__":
    main()
