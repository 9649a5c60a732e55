"""Where one `sweep.run_movie` call spends its wall clock on one GPU (bench.py's `movie_e2e` workload: an 18 000-frame movie,
179 stage-1 windows of 100 frames, stage-2 top-100 with zooms 4 / 2 / 1), with the device synchronised at the phase
boundaries, for two sizes of the batched stage-2 generate()."""
import os, sys, time
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_movie
from revisionllm_b200 import sweep, synthetic as syn
from revisionllm_b200.clip_encoder import ClipEncoder
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM

cfg = syn.VICUNA_7B_VIS
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
model.clip_encoder = ClipEncoder(model.engine, syn.make_clip_encoder_weights(cfg.hidden, seed=0, device="cuda"))
movie = synthetic_movie(cfg, 18000, seed=31).numpy()
ids = syn.make_prompt_ids(cfg, seed=2)
ids_s2 = syn.make_prompt_ids(cfg, seed=9)
cls_host = torch.randn(cfg.adapter_dim, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
q_feats = (torch.randn(1, 32, cfg.adapter_dim, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16), torch.ones(1, 32))
for per_batch in (16, 64):
    mc = sweep.MovieConfig(clip_length=200, num_frames=100, stage2_clip_length=200, stage2_num_frames=250, stride=5, batch=100,
                           zooms=(4, 2, 1), max_new_tokens=16, stage2_calls_per_batch=per_batch)
    call = lambda t=None: sweep.run_movie(model, movie, ids, cls_host, partial(syn.synthetic_answers, n_frames=100), (0.40, 0.45), mc,
                                          query_feats=q_feats, stage2_input_ids=ids_s2, detok_stage2=syn.synthetic_answers_stage2,
                                          eos_token_id=None, timings=t)
    for _ in range(3):
        res = call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    total = 1e3 * (time.perf_counter() - t0) / 2
    ph = {}
    call(ph)
    print(f"stage-2 chunks per generate() <= {per_batch}: {total:.1f} ms per movie ({len(res.stage2)} stage-2 calls over {len(res.grounding_windows)} windows); "
          + ", ".join(f"{k} {v:.1f}" for k, v in ph.items()), flush=True)
