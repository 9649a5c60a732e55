"""tcgen05 attention against the mma.sync kernels, interleaved in one process, for the two uses added late in round 2:
  (1) `rvl_mha96` (ClipEncoder attention, 8 heads x 96): self-attention over 251 tokens and text -> video cross-attention
      (250 queries x 32 text keys) for 100 and 304 windows - RVL_ATTN_MHA96 unset / 0;
  (2) the stage-1 sweep with `share_prefix_compute` (180 segments, 32 shared positions): the tcgen05 prefill kernel following
      the external context (RVL_ATTN_PREFILL unset) against the mma.sync kernel for those sequences (RVL_ATTN_PREFILL=2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import sweep, synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM

cfg = syn.VICUNA_7B_VIS
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
eng = model.engine


def set_env(name, v):
    if v is None:
        os.environ.pop(name, None)
    else:
        os.environ[name] = v
    eng.lib.rvl_reload_env()


def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ---- (1) mha96
D, nh = 768, 8
for V in (100, 304):
    for name, Tq, Tk, n_kv in (("self 251x251", 251, 251, V), ("cross 250x32", 250, 32, 1)):
        g = torch.Generator(device="cuda").manual_seed(V + Tk)
        qk = (torch.randn(V * Tq, 2 * D, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
        kv = (torch.randn(n_kv * Tk, 2 * D, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
        out = torch.empty(V * Tq, D, dtype=torch.bfloat16, device="cuda")
        idx = torch.zeros(V, dtype=torch.int32, device="cuda") if n_kv == 1 else None
        mask = torch.ones(n_kv, Tk, device="cuda") if n_kv == 1 else None
        times, outs = {}, {}
        for rep in range(7):
            for kern, env in (("tcgen05", None), ("mma.sync", "0")):
                set_env("RVL_ATTN_MHA96", env)
                t = timed(lambda: eng.mha96(qk[:, :D], kv[:, :D], kv[:, D:], out, V, nh, Tq, Tk, kv_seq_idx=idx, key_mask=mask), 8)
                if rep >= 2:
                    times.setdefault(kern, []).append(t * 1e3)
                outs[kern] = out.float().clone()
        set_env("RVL_ATTN_MHA96", None)
        med = {k: sorted(v)[len(v) // 2] for k, v in times.items()}
        diff = float((outs["tcgen05"] - outs["mma.sync"]).abs().max() / outs["mma.sync"].abs().max())
        flops = 4.0 * V * nh * Tq * Tk * 96
        print(f"mha96 {name:13s} V={V:3d}: " + "  ".join(f"{k} {med[k]:7.1f} us ({flops / med[k] / 1e6:5.1f} TFLOP/s)" for k in med)
              + f"  speed-up {med['mma.sync'] / med['tcgen05']:.2f}x  max diff {diff:.1e}", flush=True)

if os.environ.get("MHA96_ONLY"):          # ncu capture of the ClipEncoder attention alone
    sys.exit(0)

# ---- (2) shared-prefix sweep
feats = syn.make_features(180, 100, 768, seed=1, class_cfg=cfg).cuda()
ids = syn.make_prompt_ids(cfg, seed=2).cuda()
cls = torch.randn(768, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()
run = lambda: sweep.score_segments(model, feats, ids, cls, 16, eos_token_id=None)
base = run()
for _ in range(3):
    run()
t_full = [timed(run, 2) for _ in range(3)]
model.share_prefix_compute = True
times, recs = {}, {}
for rep in range(6):
    for kern, env in (("tcgen05 follows the context", None), ("mma.sync for context sequences", "2")):
        set_env("RVL_ATTN_PREFILL", env)
        if rep < 2:
            recs[kern] = run()
        else:
            times.setdefault(kern, []).append(timed(run, 2))
set_env("RVL_ATTN_PREFILL", None)
model.share_prefix_compute = False
print(f"sweep 180 x 100 frames, full computation: {sorted(t_full)[1]:.1f} ms; shared positions {model.last_shared_prefix}")
for k, v in times.items():
    same = torch.equal(sweep.unpack_records(recs[k])["tokens"], sweep.unpack_records(base)["tokens"])
    print(f"  shared prefix, {k:32s}: {sorted(v)[len(v) // 2]:.1f} ms per sweep, tokens identical to the full computation: {same}", flush=True)
