"""Q queries on one movie (180 segments x 100 frames, 7B shape, 16 greedy tokens): one sweep per query (what the reference's
evaluation loop does, eval_nlq_negative.py:183-337 once per query) against sweep.score_segments_queries with the visual
context of every segment shared by its Q prompts (`share_prefix_compute`), for several segment-batch sizes.

    python tools/multi_query.py --queries 8 --batches 32,90,180
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import sweep, synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=8)
ap.add_argument("--segments", type=int, default=180)
ap.add_argument("--batches", default="32,90,180")
args = ap.parse_args()
cfg = syn.VICUNA_7B_VIS
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
S, Q = args.segments, args.queries
feats = syn.make_features(S, 100, 768, seed=1, class_cfg=cfg).cuda()
base = syn.make_prompt_ids(cfg, seed=2)
n_pre = int((base == -200).nonzero()[0])
ids = base[None].repeat(Q, 1)
for q in range(1, Q):                                   # same system text, different query text behind <video>
    ids[q, n_pre + 1:-4] = torch.randint(3, cfg.vocab, (ids.shape[1] - n_pre - 5,), generator=torch.Generator().manual_seed(500 + q))
cls = torch.randn(Q, 768, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()


def timed(fn, reps=2):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps, out


def per_query():
    return [sweep.score_segments(model, feats, ids[q], cls[q], 16, eos_token_id=None) for q in range(Q)]


ms_single, recs_single = timed(per_query)
tok_single = torch.stack([sweep.unpack_records(r)["tokens"] for r in recs_single], dim=1).reshape(S * Q, -1)      # row = segment * Q + query
print(f"{Q} queries x {S} segments, one sweep per query: {ms_single:.0f} ms = {ms_single / Q:.1f} ms per query, "
      f"{S * Q / ms_single * 1e3:.0f} segment-queries/s", flush=True)
model.share_prefix_compute = True
for bs in [int(b) for b in args.batches.split(",")]:
    try:
        ms, rec = timed(lambda: sweep.score_segments_queries(model, feats, ids, cls, 16, batch_segments=bs, eos_token_id=None))
        same = float((sweep.unpack_records(rec)["tokens"] == tok_single).all(dim=1).float().mean())
        print(f"  all queries in one pass, shared visual context ({model.last_shared_prefix} positions), {bs} segments = {bs * Q} rows per generate(): "
              f"{ms:.0f} ms = {ms / Q:.1f} ms per query, {S * Q / ms * 1e3:.0f} segment-queries/s ({ms_single / ms:.2f}x); "
              f"rows with the same tokens as the per-query sweeps: {same:.3f}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.0f} GiB", flush=True)
    except Exception as e:
        print(f"  {bs} segments per generate(): {type(e).__name__}: {str(e)[:160]}", flush=True)
        torch.cuda.empty_cache()
model.share_prefix_compute = False
