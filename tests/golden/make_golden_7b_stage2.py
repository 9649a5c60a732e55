"""tests/golden/stage2_7b.npz: the fp32 CPU oracle's answer for three stage-2 prompts at the full shape (BASELINE.json configs[3]:
100 windows x 250 frames through the ClipEncoder adapter -> 100 CLS tokens spliced into a Vicuna-7B prompt, 16 greedy tokens) -
one prompt per zoom level of eval_nlq_retrieval_e2e2.py:337-353 (100 distinct windows; 50 windows twice; 25 windows four times).

Needs a GPU box for the Llama weights (made by the CUDA generator, exactly as bench.py and make_golden_7b.py make them):

    gpurun -- 'python tests/golden/make_golden_7b_stage2.py --out gpurun_out/stage2_7b.npz'     # then copy it to tests/golden/

The adapter weights and all inputs come from CPU generators.  Stored: the oracle's CLS rows of the first 16 windows, and per prompt
and step its greedy token, top-8 logits, the logits at 64 probe ids and the row's largest |logit|."""
import argparse
import importlib.util
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import clip_encoder_ref, llama_ref, splice_ref  # noqa: E402
from revisionllm_b200 import synthetic as syn  # noqa: E402

V, T, LQ, STEPS, N_PROBE, TOPK, N_CLS = 100, 250, 32, 16, 64, 8, 16


def _stage1_module():
    spec = importlib.util.spec_from_file_location("make_golden_7b", os.path.join(HERE, "make_golden_7b.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload(device="cuda"):
    """(cfg, llama state dict on `device`, adapter weights (CPU), windows [V, T, 768], (query tokens [1, Lq, 768], mask [1, Lq]),
    prompt ids, images [3, V, T, 768]: one stacked prompt per zoom level)."""
    cfg = syn.VICUNA_7B_VIS
    sd = syn.make_llama_weights(cfg, seed=0, device=device)
    cw = syn.make_clip_encoder_weights(cfg.hidden, seed=0)
    wins = syn.make_features(V, T, cfg.adapter_dim, seed=41)
    g = torch.Generator().manual_seed(42)
    q_tok = torch.randn(1, LQ, cfg.adapter_dim, generator=g).to(torch.bfloat16)
    q_mask = torch.ones(1, LQ)
    q_mask[0, LQ - 5:] = 0                                                 # a padded query: the key-padding mask is live
    ids = syn.make_prompt_ids(cfg, seed=9)
    rows = [torch.arange(V), torch.arange(V // 2).repeat_interleave(2), torch.arange(V // 4).repeat_interleave(4)]
    images = torch.stack([wins[r] for r in rows])
    return cfg, sd, cw, wins, (q_tok, q_mask), ids, images


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "stage2_7b.npz"))
    args = ap.parse_args()
    cfg, sd, cw, wins, (q_tok, q_mask), ids, images = workload()
    digest = _stage1_module().cheap_digest(sd)
    w32 = {k: v.detach().cpu().float() for k, v in sd.items()}
    del sd
    torch.cuda.empty_cache()
    torch.set_num_threads(os.cpu_count() or 1)
    shape = llama_ref.LlamaShape(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate, cfg.vocab, cfg.rms_eps,
                                 cfg.rope_theta, cfg.adapter_dim)
    t0 = time.perf_counter()
    cls = clip_encoder_ref.clip_encoder_cls(cw, wins.float(), q_tok.float().expand(V, -1, -1), q_mask.expand(V, -1))     # [V, hidden]
    print(f"oracle ClipEncoder over {V} windows x {T} frames: {time.perf_counter() - t0:.0f} s", flush=True)
    probe = torch.from_numpy(np.random.default_rng(19).choice(cfg.vocab, N_PROBE, replace=False).astype(np.int64))
    n = images.shape[0]
    out = dict(probe_ids=probe.numpy().astype(np.int32), tokens=np.zeros((n, STEPS), np.int32), top_ids=np.zeros((n, STEPS, TOPK), np.int32),
               top_vals=np.zeros((n, STEPS, TOPK), np.float32), probe_vals=np.zeros((n, STEPS, N_PROBE), np.float32),
               row_absmax=np.zeros((n, STEPS), np.float32), cls_rows=cls[:N_CLS].numpy(), cls_absmax=np.float32(cls.abs().max()))
    rows = [torch.arange(V), torch.arange(V // 2).repeat_interleave(2), torch.arange(V // 4).repeat_interleave(4)]
    for i, r in enumerate(rows):
        x = torch.stack(splice_ref.splice(w32, ids[None], cls[r][None]))
        toks, scores = llama_ref.greedy_decode(w32, shape, x, STEPS, stop_on_eos=False)
        sc = torch.stack(scores)[:, 0]
        top = sc.topk(TOPK, dim=-1)
        out["tokens"][i] = toks[0].numpy()
        out["top_ids"][i], out["top_vals"][i] = top.indices.numpy(), top.values.numpy()
        out["probe_vals"][i] = sc[:, probe].numpy()
        out["row_absmax"][i] = sc.abs().max(dim=-1).values.numpy()
        print(f"prompt {i} (zoom {(1, 2, 4)[i]}): tokens {toks[0].tolist()[:6]}... top-2 margin min {float((top.values[:, 0] - top.values[:, 1]).min()):.3f} "
              f"({time.perf_counter() - t0:.0f} s)", flush=True)
    out["weights_digest"] = np.array(digest)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes")


if __name__ == "__main__":
    main()
