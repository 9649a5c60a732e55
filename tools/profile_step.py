"""One warm sweep, then one sweep inside cudaProfilerStart/Stop - the region ncu captures with
`--profile-from-start off`.  Same workload as bench.py (180 segments x 100 frames, L=184, 16 tokens)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from revisionllm_b200 import sweep, synthetic as syn  # noqa: E402
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=180)
ap.add_argument("--new-tokens", type=int, default=16)
args = ap.parse_args()
cfg = syn.VICUNA_7B_VIS
sd = syn.make_llama_weights(cfg, seed=0, device="cuda")
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), sd).bfloat16().cuda()
feats = syn.make_features(args.segments, 100, 768, seed=1, class_cfg=cfg).cuda()
ids = syn.make_prompt_ids(cfg, seed=2)
cls = torch.randn(768, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()
sweep.score_segments(model, feats, ids, cls, args.new_tokens, eos_token_id=None)
torch.cuda.synchronize()
torch.cuda.profiler.start()
sweep.score_segments(model, feats, ids, cls, args.new_tokens, eos_token_id=None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one sweep; launches per sweep:", model.engine.launches // 2)
