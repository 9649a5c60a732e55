"""Phase split of the batched stage-2 pass (top-100 windows x 250 frames, zooms 4/2/1 = 7 prompts in one generate())."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import sweep, synthetic as syn
from revisionllm_b200.clip_encoder import ClipEncoder
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM
cfg = syn.VICUNA_7B
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
model.clip_encoder = ClipEncoder(model.engine, syn.make_clip_encoder_weights(cfg.hidden, seed=0, device="cuda"))
model.record_phase_events = True
wins = syn.make_features(100, 250, cfg.adapter_dim, seed=7).cuda()
g = torch.Generator().manual_seed(8)
q_tok = torch.randn(1, 32, cfg.adapter_dim, generator=g).to(torch.bfloat16)
q_mask = torch.ones(1, 32)
ids = syn.make_prompt_ids(cfg, seed=9)
for i in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sweep.stage2_pass(model, wins, (q_tok, q_mask), ids, grounding_windows=list(range(100)), batch=100, zooms=(4, 2, 1), max_new_tokens=16,
                      perm_seed=0, eos_token_id=None)
    torch.cuda.synchronize()
    ev = model.last_phase_events
    print(f"run {i}: wall {1e3 * (time.perf_counter() - t0):6.1f} ms | adapter+splice {ev[0].elapsed_time(ev[1]):6.2f} | prefill {ev[1].elapsed_time(ev[2]):6.2f} | "
          f"decode {ev[2].elapsed_time(ev[3]):6.2f}", flush=True)
