"""Deterministic synthetic weights / features / prompts for tests and the benchmark.

There is no network in the build or GPU containers, so neither Vicuna-7B /
CLIP checkpoints nor a sentencepiece vocabulary are available: everything on
the hot path is exercised with random-init weights of the right architecture
and N(0,1) features.

Token-identity on pure N(0, 0.02) weights is ill-posed (the 32000 logits are a
near-uniform Gaussian race whose top-2 margin is below bf16 noise, SURVEY.md
H1), so the generator can *plant* structure: `lm_head` carries, on top of its
random part, a copy of the token-embedding table shifted by a fixed
permutation, and `embed_tokens` gets unit scale so the current token stays
visible in the residual stream.  The greedy continuation is then the
permutation chain of the last prompt token with a margin far above bf16 noise,
while every logit still depends on all 32 layers' arithmetic.
All values are rounded to bf16 so the fp32 oracle and the bf16 CUDA path share
bit-identical parameters.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, asdict
from typing import Dict

import torch


@dataclass
class SynthConfig:
    hidden: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    head_dim: int = 128
    intermediate: int = 11008
    vocab: int = 32000
    adapter_dim: int = 768
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_pos: int = 4096
    # init
    init_std: float = 0.02
    resid_scaled_init: bool = True  # o_proj / down_proj std = init_std / sqrt(2 * n_layers) (GPT-2/NeoX-style)
    embed_std: float = 1.0
    norm_jitter: float = 0.1
    plant_gain: float = 1.0         # 0 disables the planted successor structure
    perm_mult: int = 7919           # successor(t) = 3 + ((t-3)*mult + add) mod (vocab-3)
    perm_add: int = 104729
    # > 0: the planted continuation also depends on the VISUAL input (see `plant_visual_classes`): features carry one of
    # `visual_classes` directions, a layer-0 attention head averages it into the residual stream, lm_head reads it
    visual_classes: int = 0
    class_amp: float = 8.0          # amplitude of the class direction in every frame of a segment (frame noise has norm ~27.7)
    class_gain: float = 1.0         # weight of the class read-out in lm_head, relative to plant_gain

    def dict(self):
        return asdict(self)


VICUNA_7B = SynthConfig()
VICUNA_7B_VIS = SynthConfig(visual_classes=4)      # the benchmark / full-size parity weights: four visually selected token chains
TINY = SynthConfig(hidden=256, n_layers=2, n_heads=2, head_dim=128, intermediate=512, vocab=512, max_pos=1024)
SMALL = SynthConfig(hidden=512, n_layers=4, n_heads=4, head_dim=128, intermediate=1024, vocab=2048, max_pos=2048)
MEDIUM = SynthConfig(hidden=1024, n_layers=8, n_heads=8, head_dim=128, intermediate=2816, vocab=8192, max_pos=2048)


def successor_table(cfg: SynthConfig) -> torch.Tensor:
    """successor[t] for t in [0, vocab): a permutation of [3, vocab) (ids 0..2 =
    unk/bos/eos map to 3 and are never produced), so greedy decode never hits EOS."""
    n = cfg.vocab - 3
    mult = cfg.perm_mult
    while math.gcd(mult, n) != 1:
        mult += 1
    t = torch.arange(cfg.vocab, dtype=torch.int64)
    succ = 3 + ((t - 3).clamp(min=0) * mult + cfg.perm_add) % n
    succ[:3] = 3
    return succ


def class_successor_table(cfg: SynthConfig) -> torch.Tensor:
    """[K, vocab] int64: successor of token t when the segment shows class k.  Class k's successors all lie in
    A_k = {v >= 3 : (v - 3) % K == k}, so the K candidate successors of a token are distinct and the class read-out of
    lm_head (a bonus for every token of A_k) picks one of them."""
    K = cfg.visual_classes
    per = (cfg.vocab - 3) // K
    mult = cfg.perm_mult
    while math.gcd(mult, per) != 1:
        mult += 1
    t = torch.arange(cfg.vocab, dtype=torch.int64)
    rows = []
    for k in range(K):
        slot = ((t - 3).clamp(min=0) * mult + cfg.perm_add + 7 * k) % per
        rows.append(3 + K * slot + k)
    return torch.stack(rows)


def class_directions(cfg: SynthConfig, seed: int = 0) -> torch.Tensor:
    """[K, adapter_dim] fp32 orthonormal feature-space directions, one per visual class (fixed by `seed`)."""
    g = torch.Generator().manual_seed(seed + 7777)
    q, _ = torch.linalg.qr(torch.randn(cfg.adapter_dim, cfg.visual_classes, generator=g))
    return q.t().contiguous()


def plant_visual_classes(cfg: SynthConfig, w: Dict[str, torch.Tensor], seed: int) -> None:
    """Make the greedy continuation depend on the visual input, through the model's own arithmetic:

      * every frame of a segment of class k carries `class_amp * c_k` (make_features(classes=...)); mm_projector maps it
        to p_k = P c_k in every visual position of the prompt;
      * in layer 0 the LAST attention head gets W_q = 0 (all scores 0: it averages its values over the causal context),
        value rows that read out the K directions p_k / |p_k|, and output columns that write K fixed random vectors r_k
        (embedding-sized) into the residual stream - so every position behind the video, at prefill and at every decode
        step, carries ~1.5 r_k for the class it saw;
      * lm_head: row v of A_k (class_successor_table) = noise + plant_gain * (sum of the embeddings of the tokens whose
        class-k successor is v) + class_gain * r_k.  For the current token a and class k the row succ_k(a) collects both
        bonuses, every other row at most one: the chain is succ_k(succ_k(...)) - four different chains for four classes,
        with a margin of the size of the planted bonus itself.
    Replaces the single successor permutation of plant_gain (the reference weights are random-init either way)."""
    K, H, V, d = cfg.visual_classes, cfg.hidden, cfg.vocab, cfg.head_dim
    dev = w["lm_head.weight"].device
    g = torch.Generator().manual_seed(seed + 4242)
    r = torch.randn(K, H, generator=g).to(dev)                                        # r_k, rms 1 like an embedding
    P = w["model.mm_projector.weight"].float()                                        # [H, adapter_dim]
    p = class_directions(cfg, seed).to(dev) @ P.t()                                    # [K, H]
    p_hat = p / p.norm(dim=1, keepdim=True)
    h0 = (cfg.n_heads - 1) * d
    wq, wv, wo = (w[f"model.layers.0.self_attn.{n}_proj.weight"] for n in ("q", "v", "o"))
    wq[h0:h0 + d] = 0
    wv[h0:h0 + d] = 0
    # value = beta_v * (p_hat . normed x): ~ beta_v * class_amp * |p| / rms(x) on a visual row of the class, ~ beta_v * N(0,1) elsewhere
    beta_v, beta_o = 0.25, 0.75
    wv[h0:h0 + K] = (beta_v * p_hat).to(wv.dtype)
    wo[:, h0:h0 + d] = 0
    wo[:, h0:h0 + K] = (beta_o * r.t()).to(wo.dtype)
    succ = class_successor_table(cfg).to(dev)                                          # [K, V]
    emb = w["model.embed_tokens.weight"].float() / cfg.embed_std
    head = torch.randn((V, H), generator=torch.Generator().manual_seed(seed + 99), dtype=torch.float32).to(dev)
    for k in range(K):
        head.index_add_(0, succ[k, 3:], cfg.plant_gain * emb[3:])                      # row succ_k(a) += embed[a]
        rows = torch.arange(3 + k, V, K, device=dev)
        head[rows] += cfg.class_gain * r[k]
    w["lm_head.weight"] = (head * cfg.init_std).to(torch.bfloat16)


def _randn(shape, std, gen, device, dtype=torch.bfloat16):
    return (torch.randn(shape, generator=gen, device=device, dtype=torch.float32) * std).to(dtype)


def make_llama_weights(cfg: SynthConfig, seed: int = 0, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """HF-named state dict (bf16) of a random-init Llama + Linear mm_projector.

    Names follow the reference's checkpoint layout:
    `model.embed_tokens.weight`, `model.layers.{i}.self_attn.{q,k,v,o}_proj.weight`,
    `model.layers.{i}.mlp.{gate,up,down}_proj.weight`, `...input_layernorm.weight`,
    `...post_attention_layernorm.weight`, `model.norm.weight`, `lm_head.weight`,
    `model.mm_projector.{weight,bias}`.
    The stream of random numbers depends on `device` (torch CPU and CUDA
    generators differ): generate once and copy when two devices must agree.
    """
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    H, I, V = cfg.hidden, cfg.intermediate, cfg.vocab
    w: Dict[str, torch.Tensor] = {}
    # Residual-branch output projections get the depth-scaled init so the residual stream does not grow like
    # sqrt(n_layers): with plain 0.02 the 32-layer stream reaches rms ~15 and drowns the token embedding the
    # planted successor structure relies on (measured on B200: planted logit 4.4 sigma vs a 4.1 sigma noise max).
    resid_std = cfg.init_std / math.sqrt(2 * cfg.n_layers) if cfg.resid_scaled_init else cfg.init_std
    w["model.embed_tokens.weight"] = _randn((V, H), cfg.embed_std, gen, device)
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        for n in ("q", "k", "v"):
            w[p + f"self_attn.{n}_proj.weight"] = _randn((H, H), cfg.init_std, gen, device)
        w[p + "self_attn.o_proj.weight"] = _randn((H, H), resid_std, gen, device)
        w[p + "mlp.gate_proj.weight"] = _randn((I, H), cfg.init_std, gen, device)
        w[p + "mlp.up_proj.weight"] = _randn((I, H), cfg.init_std, gen, device)
        w[p + "mlp.down_proj.weight"] = _randn((H, I), resid_std, gen, device)
        w[p + "input_layernorm.weight"] = (1.0 + _randn((H,), cfg.norm_jitter, gen, device, torch.float32)).to(torch.bfloat16)
        w[p + "post_attention_layernorm.weight"] = (1.0 + _randn((H,), cfg.norm_jitter, gen, device, torch.float32)).to(torch.bfloat16)
    w["model.norm.weight"] = (1.0 + _randn((H,), cfg.norm_jitter, gen, device, torch.float32)).to(torch.bfloat16)
    head = torch.randn((V, H), generator=gen, device=device, dtype=torch.float32)
    if cfg.plant_gain != 0.0:
        succ = successor_table(cfg).to(device)
        src = torch.zeros(V, dtype=torch.int64, device=device)
        src[succ[3:]] = torch.arange(3, V, device=device)        # src[j] = token whose successor is j
        planted = w["model.embed_tokens.weight"].float()[src] / cfg.embed_std
        planted[:3] = 0
        head = head + cfg.plant_gain * planted
    w["lm_head.weight"] = (head * cfg.init_std).to(torch.bfloat16)
    w["model.mm_projector.weight"] = _randn((H, cfg.adapter_dim), 1.0 / math.sqrt(cfg.adapter_dim), gen, device)
    w["model.mm_projector.bias"] = _randn((H,), 0.1, gen, device)
    if cfg.visual_classes > 0:
        plant_visual_classes(cfg, w, seed)
    return w


CLIP_D, CLIP_FFN, CLIP_LAYERS = 768, 2048, 2


def make_clip_encoder_weights(hidden: int, seed: int = 0, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """State dict of the stage-2 `ClipEncoder` adapter, keys below
    `model.mm_projector.` (reference module tree:
    /root/reference/revisionllm/model/adapter/transformer.py:60-92)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1000003)
    d = CLIP_D
    p: Dict[str, torch.Tensor] = {}
    p["global_rep_token"] = _randn((d,), 1.0, gen, device)
    p["global_rep_pos"] = _randn((d,), 1.0, gen, device)

    def layer(pre):
        p[pre + "self_attn.in_proj_weight"] = _randn((3 * d, d), 1.0 / math.sqrt(d), gen, device)
        p[pre + "self_attn.in_proj_bias"] = _randn((3 * d,), 0.02, gen, device)
        p[pre + "self_attn.out_proj.weight"] = _randn((d, d), 1.0 / math.sqrt(d), gen, device)
        p[pre + "self_attn.out_proj.bias"] = _randn((d,), 0.02, gen, device)
        p[pre + "linear1.weight"] = _randn((CLIP_FFN, d), 1.0 / math.sqrt(d), gen, device)
        p[pre + "linear1.bias"] = _randn((CLIP_FFN,), 0.02, gen, device)
        p[pre + "linear2.weight"] = _randn((d, CLIP_FFN), 1.0 / math.sqrt(CLIP_FFN), gen, device)
        p[pre + "linear2.bias"] = _randn((d,), 0.02, gen, device)
        for n in ("norm1", "norm2"):
            p[pre + n + ".weight"] = (1.0 + _randn((d,), 0.1, gen, device, torch.float32)).to(torch.bfloat16)
            p[pre + n + ".bias"] = _randn((d,), 0.05, gen, device)

    for i in range(CLIP_LAYERS):
        layer(f"t2v_encoder.layers.{i}.")
    for i in range(CLIP_LAYERS):
        layer(f"encoder.layers.{i}.")
    p["mm_projector.weight"] = _randn((hidden, d), 1.0 / math.sqrt(d), gen, device)
    p["mm_projector.bias"] = _randn((hidden,), 0.1, gen, device)
    return p


def make_features(n_segments: int, n_frames: int, dim: int = 768, seed: int = 0, device: str = "cpu",
                  class_cfg: "SynthConfig" = None, weight_seed: int = 0) -> torch.Tensor:
    """Synthetic CLIP frame features ~ N(0,1), bf16.  With `class_cfg` (a config with
    visual_classes > 0) every frame of segment i additionally carries the direction of class `segment_class(i, K)`
    (plant_visual_classes: the weights made with `weight_seed` read it back out)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 17)
    x = torch.randn((n_segments, n_frames, dim), generator=gen, device=device, dtype=torch.float32)
    if class_cfg is not None and class_cfg.visual_classes > 0:
        dirs = class_directions(class_cfg, weight_seed).to(device)
        cls = segment_classes(n_segments, class_cfg.visual_classes).to(device)
        x = x + class_cfg.class_amp * dirs[cls][:, None, :]
    return x.to(torch.bfloat16)


def segment_classes(n_segments: int, n_classes: int) -> torch.Tensor:
    """Visual class of each synthetic segment: a fixed pseudo-random pattern (not i % K, so neighbours differ irregularly)."""
    i = torch.arange(n_segments, dtype=torch.int64)
    return ((i * 2654435761) >> 7) % n_classes


def expected_chain(cfg: SynthConfig, last_prompt_token: int, steps: int, visual_class: int = 0) -> list:
    """The planted greedy continuation: successor chain of the last prompt token (of class `visual_class` when the weights
    carry visual classes)."""
    table = class_successor_table(cfg)[visual_class] if cfg.visual_classes > 0 else successor_table(cfg)
    out, cur = [], int(last_prompt_token)
    for _ in range(steps):
        cur = int(table[cur])
        out.append(cur)
    return out


def make_prompt_ids(cfg: SynthConfig, n_pre: int = 38, n_post: int = 46, seed: int = 0) -> torch.Tensor:
    """BOS + (n_pre-1) text ids + IMAGE placeholder (-200) + n_post text ids.
    Defaults give Ltxt = 85 and, with 100 frames, a spliced length of 184
    (BASELINE.md section 4)."""
    gen = torch.Generator()
    gen.manual_seed(seed + 29)
    pre = torch.randint(3, cfg.vocab, (n_pre - 1,), generator=gen)
    post = torch.randint(3, cfg.vocab, (n_post,), generator=gen)
    return torch.cat([torch.tensor([1]), pre, torch.tensor([-200]), post]).to(torch.int64)


def weights_digest(w: Dict[str, torch.Tensor]) -> str:
    """sha256 over names + raw bytes - fixtures store it to detect generator drift."""
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        t = w[k].detach().cpu().contiguous()
        h.update(t.view(torch.uint8).numpy().tobytes() if t.dtype != torch.bfloat16 else t.view(torch.int16).numpy().tobytes())
    return h.hexdigest()


class StubTokenizer:
    """Minimal whitespace tokenizer with the sentencepiece-Llama surface the
    reference's `inference()` touches (`__call__().input_ids`, `bos_token_id`,
    `batch_decode`).  Used only because no real vocabulary is available offline.
    Words are hashed into [3 + n_special, vocab); digits and the answer words
    have fixed ids so `"From 12 to 34."` / `"Not Present"` round-trip."""

    bos_token_id = 1
    eos_token_id = 2
    pad_token_id = 0
    SPECIAL = ["From", "to", "Not", "Present", ".", "and"] + [str(d) for d in range(10)]

    def __init__(self, vocab: int = 32000):
        self.vocab = vocab
        self.fixed = {s: 3 + i for i, s in enumerate(self.SPECIAL)}
        self.inv = {v: k for k, v in self.fixed.items()}

    def _word(self, wd: str) -> int:
        if wd in self.fixed:
            return self.fixed[wd]
        hv = int(hashlib.md5(wd.encode()).hexdigest()[:8], 16)
        base = 3 + len(self.SPECIAL)
        return base + hv % (self.vocab - base)

    def __call__(self, text: str):
        ids = [self.bos_token_id]
        for wd in text.replace("\n", " \n ").split(" "):
            if wd == "":
                continue
            if wd.isdigit():
                ids.extend(self.fixed[c] for c in wd)
            else:
                ids.append(self._word(wd))

        class _R:
            pass

        r = _R()
        r.input_ids = ids
        return r

    def batch_decode(self, ids, skip_special_tokens: bool = True):
        out = []
        for row in ids.tolist():
            s, prev_digit = "", False
            for t in row:
                if skip_special_tokens and t in (0, 1, 2):
                    continue
                tok = self.inv.get(t, f"<{t}>")
                if tok.isdigit() and prev_digit:
                    s += tok
                elif tok == ".":
                    s += tok
                else:
                    s += (" " if s else "") + tok
                prev_digit = tok.isdigit()
            out.append(s)
        return out


def synthetic_answers(tokens, n_frames: int):
    """Stand-in for `tokenizer.batch_decode` on random-init weights (no vocabulary offline): a deterministic map from a row of
    generated token ids to one of the answer strings the fine-tuned model produces - "Not Present" or "From a to b." with a
    span inside the window - so that the parse -> select -> stage-2 -> rank chain has something to chew on."""
    out = []
    for row in (tokens.tolist() if hasattr(tokens, "tolist") else tokens):
        if row[0] % 5 == 0:
            out.append("Not Present")
        else:
            a = row[1 % len(row)] % (n_frames - 1)           # never the "From F-1 to F-1." sentinel the drivers drop
            b = min(n_frames - 1, a + row[2 % len(row)] % 23)
            out.append(f"From {a} to {b}.")
    return out


def synthetic_answers_stage2(tokens):
    """Stage-2 answers name a (zoomed, permuted) window number: here the first generated id modulo 97."""
    return [str(row[0] % 97) for row in (tokens.tolist() if hasattr(tokens, "tolist") else tokens)]
