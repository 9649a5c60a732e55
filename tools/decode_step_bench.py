"""Decode-step timing at the 7B shape: prefill B segments of L=184, then time decode steps with CUDA events.
Reports ms/step against the HBM roofline (13.214 GB of weights + KV reads per step).  Env switches (RVL_PDL, ...) are
read by the library, so A/B runs are separate processes."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from revisionllm_b200 import synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="180,1")
ap.add_argument("--steps", type=int, default=15)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
cfg = syn.VICUNA_7B
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
eng = model.engine
ids = syn.make_prompt_ids(cfg, seed=2)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6531.9
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("RVL_"))
AB = os.environ.get("AB_ENV")      # e.g. AB_ENV=RVL_SPAIR (odd repetitions run with it set to 0) or AB_ENV=RVL_PDL:5; read per call by the library
AB_VAL = "0"
if AB and ":" in AB:
    AB, AB_VAL = AB.split(":")
for B in [int(b) for b in args.batches.split(",")]:
    feats = syn.make_features(B, 100, 768, seed=1).cuda()
    best = None
    ab = {0: [], 1: []}
    for rep in range(args.reps if not AB else 2 * args.reps + 1):
        if AB:
            if rep % 2:
                os.environ[AB] = AB_VAL
            else:
                os.environ.pop(AB, None)
            eng.lib.rvl_reload_env()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        out = model(ids[None].expand(B, -1), images=feats, logits_to_keep=1, reserve_new_tokens=64)
        kv = out.past_key_values
        logits = out.logits[:, 0].contiguous()
        tok = torch.empty(B, dtype=torch.int32, device="cuda")
        ent = torch.empty(B, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        e[0].record()
        for t in range(args.steps):
            eng.sample_greedy(logits, tok, ent, None, -1, 0)
            eng.decode_step(tok, kv.seq_lens, kv.page_table, logits, max_kv_len=int(kv.lengths.max()) + t + 1)
        e[1].record()
        host_ms = 1e3 * (time.perf_counter() - t0) / args.steps          # host enqueue time per step (the queue never fills in 15 steps)
        torch.cuda.synchronize()
        ms = e[0].elapsed_time(e[1]) / args.steps
        best = ms if best is None else min(best, ms)
        if AB and rep > 0:
            ab[rep % 2].append(ms)
    if AB:
        print(f"B={B}: default {sum(ab[0]) / len(ab[0]):.3f} ms (min {min(ab[0]):.3f}) | {AB}={AB_VAL} {sum(ab[1]) / len(ab[1]):.3f} ms (min {min(ab[1]):.3f})", flush=True)
        os.environ.pop(AB, None)
        continue
    L = ids.shape[0] - 1 + 100
    bytes_step = 13.214e9 + B * 0.524288e6 * (L + args.steps / 2 + 1)
    print(f"[{tag}] B={B:4d}: {best:7.3f} ms/decode step   {bytes_step / best / 1e6:7.0f} GB/s = {bytes_step / best / 1e6 / peak:.3f} of HBM peak   (host enqueue {host_ms:.3f} ms/step)", flush=True)
