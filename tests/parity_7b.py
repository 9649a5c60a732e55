"""Full-size parity evidence (north star: tokens identical on >= 99 % of segments, logits within a stated bf16 tolerance):
N segments of the 1-hour sweep at the Vicuna-7B shape, CUDA path (batched) vs the fp32 CPU oracle (one segment at a time).
Writes one JSON line; run on the GPU box:  python tests/parity_7b.py 8 > gpurun_out/parity_7b.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import torch
from oracle import llama_ref, splice_ref
from revisionllm_b200 import synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = 16
cfg = syn.VICUNA_7B
sd = syn.make_llama_weights(cfg, seed=0, device="cuda")
w32 = {k: v.detach().cpu().float() for k, v in sd.items()}
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), sd).bfloat16().cuda()
feats = syn.make_features(180, 100, cfg.adapter_dim, seed=1)
ids = syn.make_prompt_ids(cfg, seed=2)
pick = list(range(0, 180, 180 // n))[:n]
out = model.generate(ids[None].repeat(180, 1), images=feats, max_new_tokens=steps, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
got_tok = out["sequences"][:, ids.shape[0]:].cpu()
got_sc = torch.stack(out["scores"]).cpu()                       # [steps, 180, V]
shape = llama_ref.LlamaShape(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate, cfg.vocab, cfg.rms_eps, cfg.rope_theta, cfg.adapter_dim)
torch.set_num_threads(os.cpu_count() or 1)
same, errs, t0 = 0, [], time.perf_counter()
for i in pick:
    x = torch.stack(splice_ref.splice(w32, ids[None], splice_ref.mm_projector_linear(w32, feats[i:i + 1].float())))
    toks, scores = llama_ref.greedy_decode(w32, shape, x, steps, stop_on_eos=False)
    ref = torch.stack(scores)[:, 0]
    same += int(torch.equal(toks[0].long(), got_tok[i].long()))
    errs.append(float((got_sc[:, i].double() - ref.double()).abs().max() / ref.double().abs().max()))
print(json.dumps({"segments_checked": len(pick), "segments": pick, "batch": 180, "new_tokens": steps, "tokens_identical": same,
                  "identical_fraction": same / len(pick), "logit_max_rel_err": max(errs), "logit_rel_err_per_segment": errs,
                  "tolerance": 3e-2, "oracle_seconds": time.perf_counter() - t0, "oracle_threads": torch.get_num_threads()}))
