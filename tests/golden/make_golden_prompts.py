"""More placeholder-tokenisation cases from the reference's own `tokenizer_image_token` (run where /root/reference exists):

    python tests/golden/make_golden_prompts.py

Imports /root/reference/revisionllm/mm_utils.py through oracle/ref_shim.py and runs it with the stub tokenizer (with and
without a BOS token) on prompts with one and several <video> placeholders, a <memory> placeholder with and without text
behind it, an empty chunk in front of the placeholder.  Output: tests/golden/prompt_more.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from revisionllm_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class NoBos(syn.StubTokenizer):
    """A tokenizer that does not prepend BOS (offset 0 in mm_utils.py:48-51)."""
    def __call__(self, text):
        r = super().__call__(text)
        r.input_ids = r.input_ids[1:]
        return r


def main():
    ref = ref_shim.load()
    prompts = [
        "A chat. USER: <video>\nWhat happens at 12 ? ASSISTANT:",
        "<video>\nDuring which frames can we see a dog ?",
        "USER: <video> and <video>\ncompare them ASSISTANT:",
        "USER: <video>\nWhere is the cat ?<memory> ASSISTANT:",
        "USER: <video>\nWhere is the cat ?<memory>",
        "USER: first <video> then <video> then <video> end",
    ]
    cases = []
    for bos in (True, False):
        tok = syn.StubTokenizer(32000) if bos else NoBos(32000)
        for p in prompts:
            ids = ref.mm_utils.tokenizer_image_token(p, tok, ref.constants.IMAGE_TOKEN_INDEX, return_tensors="pt")
            cases.append(dict(prompt=p, bos=bos, ids=[int(v) for v in ids]))
    json.dump(dict(source="revisionllm/mm_utils.py:22-75", cases=cases), open(os.path.join(HERE, "prompt_more.json"), "w"))
    print(len(cases), "cases;", [len(c["ids"]) for c in cases])


if __name__ == "__main__":
    main()
