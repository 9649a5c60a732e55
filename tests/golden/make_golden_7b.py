"""tests/golden/stage1_7b_vis.npz: the fp32 CPU oracle's answer for 30 of the 180 segments of the benchmark workload at the
full Vicuna-7B shape (BASELINE.json configs[1]: L = 184, 16 greedy tokens), so that the `-m gpu` suite and bench.py can
check the CUDA path against the oracle at the BASELINE shape without spending minutes of CPU time per run.

Needs a GPU box (the benchmark's weights come from the CUDA generator, exactly as bench.py makes them) and ~60 GB of host RAM:

    gpurun -- 'python tests/golden/make_golden_7b.py --out gpurun_out/stage1_7b_vis.npz'   # then copy it to tests/golden/

Stored per checked segment and step: the oracle's greedy token, its top-8 logits (values + ids), the logits at 64 fixed
probe ids, the row's largest |logit| and the top-2 margin.  The oracle itself (oracle/llama_ref.py, oracle/splice_ref.py) is
pinned to the reference by the fixtures of make_golden.py; this file only caches its output."""
import argparse
import hashlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import llama_ref, splice_ref  # noqa: E402
from revisionllm_b200 import synthetic as syn  # noqa: E402

N_SEG, N_FRAMES, STEPS, N_CHECK, N_PROBE, TOPK = 180, 100, 16, 30, 64, 8


def cheap_digest(sd) -> str:
    """sha256 over a few slices of the weights (the full 13 GB would take longer than the test that reads this)."""
    h = hashlib.sha256()
    for name in ("model.embed_tokens.weight", "lm_head.weight", "model.layers.0.self_attn.v_proj.weight",
                 "model.layers.31.mlp.down_proj.weight", "model.mm_projector.weight"):
        t = sd[name][:64].detach().cpu().contiguous().view(torch.int16)
        h.update(name.encode())
        h.update(t.numpy().tobytes())
    return h.hexdigest()


def workload(device="cuda"):
    cfg = syn.VICUNA_7B_VIS
    sd = syn.make_llama_weights(cfg, seed=0, device=device)
    feats = syn.make_features(N_SEG, N_FRAMES, cfg.adapter_dim, seed=1, class_cfg=cfg)
    ids = syn.make_prompt_ids(cfg, seed=2)
    return cfg, sd, feats, ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage1_7b_vis.npz"))
    args = ap.parse_args()
    cfg, sd, feats, ids = workload()
    digest = cheap_digest(sd)
    w32 = {k: v.detach().cpu().float() for k, v in sd.items()}
    del sd
    torch.cuda.empty_cache()
    torch.set_num_threads(os.cpu_count() or 1)
    shape = llama_ref.LlamaShape(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate, cfg.vocab, cfg.rms_eps,
                                 cfg.rope_theta, cfg.adapter_dim)
    segs = list(range(0, N_SEG, N_SEG // N_CHECK))[:N_CHECK]
    probe = torch.from_numpy(np.random.default_rng(9).choice(cfg.vocab, N_PROBE, replace=False).astype(np.int64))
    out = dict(segments=np.array(segs, np.int32), probe_ids=probe.numpy().astype(np.int32), tokens=np.zeros((N_CHECK, STEPS), np.int32),
               top_ids=np.zeros((N_CHECK, STEPS, TOPK), np.int32), top_vals=np.zeros((N_CHECK, STEPS, TOPK), np.float32),
               probe_vals=np.zeros((N_CHECK, STEPS, N_PROBE), np.float32), row_absmax=np.zeros((N_CHECK, STEPS), np.float32),
               classes=syn.segment_classes(N_SEG, cfg.visual_classes).numpy().astype(np.int32)[segs])
    t0 = time.perf_counter()
    for n, i in enumerate(segs):
        x = torch.stack(splice_ref.splice(w32, ids[None], splice_ref.mm_projector_linear(w32, feats[i:i + 1].float())))
        toks, scores = llama_ref.greedy_decode(w32, shape, x, STEPS, stop_on_eos=False)
        sc = torch.stack(scores)[:, 0]                                   # [steps, V]
        top = sc.topk(TOPK, dim=-1)
        out["tokens"][n] = toks[0].numpy()
        out["top_ids"][n], out["top_vals"][n] = top.indices.numpy(), top.values.numpy()
        out["probe_vals"][n] = sc[:, probe].numpy()
        out["row_absmax"][n] = sc.abs().max(dim=-1).values.numpy()
        print(f"segment {i} class {out['classes'][n]} tokens {toks[0].tolist()[:6]}... margin min {float((top.values[:, 0] - top.values[:, 1]).min()):.3f} "
              f"({time.perf_counter() - t0:.0f} s)", flush=True)
    out["weights_digest"] = np.array(digest)
    out["config"] = np.array(repr(cfg))
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes")


if __name__ == "__main__":
    main()
