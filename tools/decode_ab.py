"""Interleaved A/B of decode-step variants at the 7B shape inside ONE process (the power-capped clock drifts between
processes): every variant = a set of RVL_* switches (re-read through rvl_reload_env) x {eager launches, CUDA-graph replay}.
Prefill B segments of L = 184 once per batch size, then time chunks of 8 decode steps (engine.decode_chunk) with CUDA events.

    python tools/decode_ab.py --batches 180,23 --reps 5 --variants "base;RVL_SPAIR_SMALL=1;RVL_SPAIR_SMALL=1,RVL_PDL=13"
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="180,23")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--variants", default="base;RVL_SPAIR_SMALL=1;RVL_SPAIR_SMALL=1,RVL_PDL=13;RVL_PDL=13")
ap.add_argument("--graph", default="0,1")
ap.add_argument("--out", default="")
args = ap.parse_args()
cfg = syn.VICUNA_7B
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
eng = model.engine
ids = syn.make_prompt_ids(cfg, seed=2)
peak = 6531.9
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
variants = []
for v in args.variants.split(";"):
    env = {} if v == "base" else dict(kv.split("=") for kv in v.split(","))
    for g in [int(x) for x in args.graph.split(",")]:
        variants.append((v + ("+graph" if g else ""), env, bool(g)))


def set_env(env):
    for k in [k for k in os.environ if k.startswith("RVL_")]:
        del os.environ[k]
    os.environ.update(env)
    eng.lib.rvl_reload_env()


results = {}
K = eng.DECODE_CHUNK
for B in [int(b) for b in args.batches.split(",")]:
    feats = syn.make_features(B, 100, 768, seed=1).cuda()
    L = ids.shape[0] - 1 + 100
    out = model(ids[None].expand(B, -1), images=feats, logits_to_keep=1, reserve_new_tokens=64)
    kv = out.past_key_values
    bufs = eng.decode_buffers(B, kv.page_table.shape[1])
    bufs["page_table"].copy_(kv.page_table)
    bufs["logits"].copy_(out.logits[:, 0])
    graphs = {name: {} for name, _, _ in variants}
    probe = torch.zeros(2, dtype=torch.int64, device="cuda")
    mhz = {}
    times = {name: [] for name, _, _ in variants}
    for rep in range(args.reps + 2):
        for name, env, g in variants:
            set_env(env)
            eng._dec_graphs = graphs[name]
            bufs["seq_lens"].copy_(kv.seq_lens)            # every run decodes positions L .. L + 15 again
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(2):
                eng.decode_chunk(bufs, K, -1, 0, False, L + 64, graph=g)
            e1.record()
            eng.lib.rvl_debug_sm_clock(probe.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            mhz[name] = 1e3 * probe[0].item() / max(probe[1].item(), 1)
            if rep >= 2:                                   # rep 0: eager warm-up, rep 1: capture
                times[name].append(e0.elapsed_time(e1) / (2 * K))
    bytes_step = 13.214e9 + B * 0.524288e6 * (L + K + 1)
    for name in times:
        t = sorted(times[name])
        med = t[len(t) // 2]
        results[f"B={B} {name}"] = {"ms_per_step_median": med, "min": t[0], "max": t[-1], "gbs": bytes_step / med / 1e6,
                                    "frac_hbm": bytes_step / med / 1e6 / peak}
        print(f"B={B:4d} {name:55s} median {med:7.3f} ms/step (min {t[0]:.3f} max {t[-1]:.3f})  {bytes_step / med / 1e6:6.0f} GB/s = {bytes_step / med / 1e6 / peak:.3f} of HBM   SM clock after the run {mhz[name]:.0f} MHz", flush=True)
set_env({})
if args.out:
    json.dump(results, open(args.out, "w"), indent=1)
