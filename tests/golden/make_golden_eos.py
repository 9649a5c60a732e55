"""Golden EOS bookkeeping produced by EXECUTING the reference's own lines (run where /root/reference exists):

    python tests/golden/make_golden_eos.py

/root/reference/revisionllm/model/vtimellm_llama.py:340-362 - from "finished sentences should have their next token be a
padding token" to the all-finished check - is cut out of the reference's copy of `sample()`, dedented and exec'd step by
step on scripted `next_tokens` (what the multinomial draw / argmax produced), with stubs for the pieces around it.
Output: tests/golden/eos_rule.json - per case the tokens appended at every step, the unfinished flags after it, and the step
at which the loop stops."""
import json
import os
import textwrap

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/revisionllm/model/vtimellm_llama.py"


class _Self:
    class config:
        is_encoder_decoder = False

    @staticmethod
    def _update_model_kwargs_for_generation(outputs, model_kwargs, is_encoder_decoder=False):
        return model_kwargs


def main():
    lines = open(SRC).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == "# finished sentences should have their next token be a padding token")
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip() == "this_peer_finished = True")
    src = textwrap.dedent("\n".join(lines[i0:i1 + 1]))
    g = torch.Generator().manual_seed(5)
    cases = []
    for B, T, eos, pad, p_eos in [(4, 8, 2, 2, 0.25), (3, 6, 2, 0, 0.3), (5, 10, 2, 2, 0.05), (1, 5, 2, 2, 0.5), (6, 12, 7, 3, 0.2)]:
        raw = torch.randint(10, 500, (T, B), generator=g)
        raw[torch.rand(T, B, generator=g) < p_eos] = eos
        ns = dict(torch=torch, self=_Self, eos_token_id=[eos], pad_token_id=pad, eos_token_id_tensor=torch.tensor([eos]),
                  unfinished_sequences=torch.ones(B, dtype=torch.long), input_ids=torch.zeros(B, 1, dtype=torch.long), streamer=None,
                  outputs=None, model_kwargs={}, this_peer_finished=False)
        appended, flags, stop = [], [], None
        for t in range(T):
            ns["next_tokens"] = raw[t].clone()
            exec(src, ns)
            appended.append(ns["input_ids"][:, -1].tolist())
            flags.append(ns["unfinished_sequences"].tolist())
            if ns["this_peer_finished"]:
                stop = t
                break
        cases.append(dict(eos=eos, pad=pad, raw=raw.tolist(), appended=appended, unfinished=flags, stopped_after_step=stop))
    json.dump(dict(source=f"vtimellm_llama.py:{i0 + 1}-{i1 + 1}", cases=cases), open(os.path.join(HERE, "eos_rule.json"), "w"))
    print(f"lines {i0 + 1}-{i1 + 1}", [(len(c["appended"]), c["stopped_after_step"]) for c in cases])


if __name__ == "__main__":
    main()
