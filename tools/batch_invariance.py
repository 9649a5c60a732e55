"""Batch-of-1 vs batch-of-6 logits at the 7B shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM
cfg = syn.VICUNA_7B
m = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
ids = syn.make_prompt_ids(cfg, seed=2)
feats = syn.make_features(6, 100, cfg.adapter_dim, seed=1)
steps = int(os.environ.get("STEPS", "8"))
def run(f):
    o = m.generate(ids[None].repeat(f.shape[0], 1), images=f, max_new_tokens=steps, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    return torch.stack(o["scores"]).float().cpu(), o["sequences"][:, ids.shape[0]:].cpu()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
s6, t6 = run(feats)
s1, t1 = run(feats[2:3])
s6b, _ = run(feats)
print("B=1 vs B=6 per step:", [round(rel(s1[t, 0], s6[t, 2]), 5) for t in range(steps)], "tokens equal", t1[0].tolist() == t6[2].tolist(),
      "| run-to-run B=6:", rel(s6b, s6))
