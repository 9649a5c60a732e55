"""Probe the planted-successor margin at the 7B shape on the GPU (prints logit statistics)."""
import os, sys, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM

for gain in (1.0, 2.0, 4.0):
    cfg = dataclasses.replace(syn.VICUNA_7B, plant_gain=gain)
    w = syn.make_llama_weights(cfg, seed=0, device="cuda")
    m = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), w).bfloat16().cuda()
    del w
    feats = syn.make_features(4, 100, 768, seed=1)
    ids = syn.make_prompt_ids(cfg, seed=2)
    out = m.generate(ids[None].repeat(4, 1), images=feats, max_new_tokens=4, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    succ = syn.successor_table(cfg)
    cur = int(ids[-1])
    for t, sc in enumerate(out["scores"]):
        cur = int(succ[cur])
        top = torch.topk(sc, 3, dim=-1)
        print(f"gain {gain} step {t}: expected {cur} logit {sc[:, cur].tolist()} top3 {top.values[0].tolist()} idx {top.indices[0].tolist()} std {float(sc.std()):.3f} finite {bool(torch.isfinite(sc).all())}")
    del m
    torch.cuda.empty_cache()
