"""Generate the committed golden fixtures from the REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules through `oracle/ref_shim.py`
(`prepare_inputs_labels_for_multimodal`, `VTimeLLMLlamaForCausalLM.forward`,
`ClipEncoder`, `_topk_pooling`, `get_entropy_statistics`,
`tokenizer_image_token`, `conv_templates['v1']`, `pad_sequences_1d`) plus the
installed `transformers` Llama (eager attention, fp32), runs them on seeded
synthetic inputs from `revisionllm_b200.synthetic`, and writes input/output
vectors to `tests/golden/*.npz`.  Weights are regenerated from the seed by the
tests; each fixture stores their sha256 so generator drift is detected.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from revisionllm_b200 import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def greedy_via_reference(model, inputs_embeds, attention_mask, steps):
    """Manual greedy loop around the reference model's forward (its
    `generate()` breaks at the first decode step under transformers 5.x,
    SURVEY.md section 8c).  Right-padded rows: the step-t position is the number of
    valid tokens so far (vtimellm_arch.py:88-100)."""
    B, L, _ = inputs_embeds.shape
    lens = attention_mask.sum(1)
    pos = (attention_mask.long().cumsum(1) - 1).clamp(min=0)
    out = model(inputs_embeds=inputs_embeds, attention_mask=attention_mask.long(), position_ids=pos, use_cache=True)
    logits_all = out.logits.float()
    last = logits_all[torch.arange(B), lens - 1]
    cache = out.past_key_values
    am = attention_mask.long()
    toks, scores = [], []
    for t in range(steps):
        scores.append(last.clone())
        nxt = last.argmax(-1)
        toks.append(nxt)
        if t == steps - 1:
            break
        am = torch.cat([am, torch.ones(B, 1, dtype=am.dtype)], dim=1)
        pid = am.sum(1, keepdim=True) - 1
        emb = model.get_model().embed_tokens(nxt)[:, None]
        out = model(inputs_embeds=emb, attention_mask=am, position_ids=pid, past_key_values=cache, use_cache=True)
        cache = out.past_key_values
        last = out.logits[:, -1].float()
    return logits_all, torch.stack(toks, 1), torch.stack(scores, 0)


def case_stage1(ref, name, cfg, n_seg, n_frames, n_pre, n_post, steps, ragged=False):
    w = syn.make_llama_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w)
    feats = syn.make_features(n_seg, n_frames, cfg.adapter_dim, seed=1).float()
    base = syn.make_prompt_ids(cfg, n_pre, n_post, seed=2)
    ids = base[None].repeat(n_seg, 1)
    attn = None
    if ragged:
        # rows get different numbers of trailing text tokens (right padded with 0)
        Ltxt = ids.shape[1]
        attn = torch.ones(n_seg, Ltxt, dtype=torch.bool)
        for b in range(n_seg):
            cut = (b * 3) % 7
            if cut:
                attn[b, Ltxt - cut:] = False
                ids[b, Ltxt - cut:] = 0
    with torch.inference_mode():
        r = model.prepare_inputs_labels_for_multimodal(
            ids, None, attn, None, None, feats, None, None, None, None)
        _, pos, am, _, embeds, _ = r
        if am is None:
            am = torch.ones(embeds.shape[:2], dtype=torch.bool)
        logits_all, toks, scores = greedy_via_reference(model, embeds, am.bool(), steps)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cfg=np.array(list(cfg.dict().items()), dtype=object), digest=syn.weights_digest(w),
        feats=feats.numpy(), ids=ids.numpy(), attn=(attn.numpy() if attn is not None else np.zeros(0)),
        embeds=embeds.numpy(), embeds_mask=am.bool().numpy(),
        prefill_logits=logits_all.numpy().astype(np.float32), tokens=toks.numpy(), scores=scores.numpy(),
    )
    print(name, "embeds", tuple(embeds.shape), "tokens", toks[0].tolist())


def case_stage1_truncated(ref, name, cfg, n_seg, n_frames, n_pre, n_post, max_len):
    """`tokenizer_model_max_length` (vtimellm_arch.py:239-243): every spliced row is cut to max_len positions - inside the
    visual block for the long rows, not at all for the short ones - then right-padded (:245-276).  Splice only."""
    w = syn.make_llama_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w)
    model.config.tokenizer_model_max_length = max_len
    feats = syn.make_features(n_seg, n_frames, cfg.adapter_dim, seed=31).float()
    base = syn.make_prompt_ids(cfg, n_pre, n_post, seed=32)
    ids = base[None].repeat(n_seg, 1)
    Ltxt = ids.shape[1]
    attn = torch.ones(n_seg, Ltxt, dtype=torch.bool)
    for b in range(n_seg):
        cut = (b * 5) % 9                        # 0 .. 8 trailing text ids masked out
        if cut:
            attn[b, Ltxt - cut:] = False
            ids[b, Ltxt - cut:] = 0
    with torch.inference_mode():
        r = model.prepare_inputs_labels_for_multimodal(ids, None, attn, None, None, feats, None, None, None, None)
        _, pos, am, _, embeds, _ = r
    np.savez_compressed(os.path.join(OUT, name + ".npz"), digest=syn.weights_digest(w), feats=feats.numpy(), ids=ids.numpy(), attn=attn.numpy(),
                        max_len=np.int64(max_len), embeds=embeds.numpy(), embeds_mask=am.bool().numpy())
    print(name, "embeds", tuple(embeds.shape), "row lengths", am.sum(1).tolist())


def case_stage1_no_placeholder(ref, name, cfg, n_seg, n_frames, n_pre, n_post, rows_without, max_len=None):
    """A batch in which some rows hold no <video> placeholder (vtimellm_arch.py:168-176): such a row is text only, but it
    still consumes its visual block (`cur_image_idx += 1`), so the following rows keep theirs.  Splice only."""
    w = syn.make_llama_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w)
    if max_len is not None:
        model.config.tokenizer_model_max_length = max_len       # :239-243 cuts the text-only rows as well
    feats = syn.make_features(n_seg, n_frames, cfg.adapter_dim, seed=41).float()
    base = syn.make_prompt_ids(cfg, n_pre, n_post, seed=42)
    ids = base[None].repeat(n_seg, 1)
    for b in rows_without:
        ids[b][ids[b] == ref.constants.IMAGE_TOKEN_INDEX] = 7 + b
    with torch.inference_mode():
        r = model.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, feats, None, None, None, None)
        _, _, am, _, embeds, _ = r
    if am is None:
        am = torch.ones(embeds.shape[:2], dtype=torch.bool)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), digest=syn.weights_digest(w), feats=feats.numpy(), ids=ids.numpy(),
                        embeds=embeds.numpy(), embeds_mask=am.bool().numpy(), max_len=np.int64(-1 if max_len is None else max_len))
    print(name, "embeds", tuple(embeds.shape), "row lengths", am.sum(1).tolist())


def case_stage1_image_list(ref, name, cfg, frames, n_pre, n_post):
    """`images` as a LIST of [F_i, 768] tensors with different F_i (vtimellm_arch.py:102-109: one projector call over the
    concatenation, split back per row), prompts right-padded with an attention mask - the VidChapters-shaped ragged batch."""
    w = syn.make_llama_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w)
    n_seg = len(frames)
    images = [syn.make_features(1, f, cfg.adapter_dim, seed=50 + i)[0].float() for i, f in enumerate(frames)]
    base = syn.make_prompt_ids(cfg, n_pre, n_post, seed=52)
    ids = base[None].repeat(n_seg, 1)
    Ltxt = ids.shape[1]
    attn = torch.ones(n_seg, Ltxt, dtype=torch.bool)
    for b in range(n_seg):
        cut = (b * 2) % 5
        if cut:
            attn[b, Ltxt - cut:] = False
            ids[b, Ltxt - cut:] = 0
    with torch.inference_mode():
        r = model.prepare_inputs_labels_for_multimodal(ids, None, attn, None, None, images, None, None, None, None)
        _, _, am, _, embeds, _ = r
    np.savez_compressed(os.path.join(OUT, name + ".npz"), digest=syn.weights_digest(w), frames=np.array(frames),
                        images=torch.cat(images).numpy(), ids=ids.numpy(), attn=attn.numpy(), embeds=embeds.numpy(), embeds_mask=am.bool().numpy())
    print(name, "embeds", tuple(embeds.shape), "row lengths", am.sum(1).tolist())


def case_two_placeholders(ref, name, cfg):
    """A row with TWO <video> placeholders consumes two visual blocks, the next row the one after (`cur_image_idx`,
    vtimellm_arch.py:178-207); `images` is a list with one entry per placeholder."""
    w = syn.make_llama_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w)
    frames = (3, 5, 2)
    images = [syn.make_features(1, f, cfg.adapter_dim, seed=60 + i)[0].float() for i, f in enumerate(frames)]
    base = syn.make_prompt_ids(cfg, 4, 8, seed=62)
    ids = base[None].repeat(2, 1)
    ids[0, 9] = ref.constants.IMAGE_TOKEN_INDEX                 # a second placeholder further down row 0
    with torch.inference_mode():
        r = model.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, images, None, None, None, None)
        embeds = r[4]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), digest=syn.weights_digest(w), frames=np.array(frames),
                        images=torch.cat(images).numpy(), ids=ids.numpy(), embeds=embeds.numpy())
    print(name, "embeds", tuple(embeds.shape))


def case_decode_fixup(ref, name, cfg):
    """The one-token branch of prepare_inputs_labels_for_multimodal (vtimellm_arch.py:88-100): the prompt-time mask is
    extended with ones up to the cache length + 1 and the position of the new token is sum(mask) - 1 - for right-padded
    rows the padding hole stays in the mask and the new token's position follows the row's own length."""
    w = syn.make_llama_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w)
    rows = {}
    for key, (lens, past_len) in dict(step1=((9, 6, 9, 4), 9), step3=((9, 6, 9, 4), 11), full=((7, 7), 7)).items():
        B, L = len(lens), max(lens)
        mask = torch.zeros(B, L, dtype=torch.long)
        for b, n in enumerate(lens):
            mask[b, :n] = 1
        if key == "step3":                       # two earlier steps already appended their ones
            mask = torch.cat([mask, torch.ones(B, 2, dtype=torch.long)], dim=1)
        past = ((torch.zeros(B, cfg.n_heads, past_len, cfg.head_dim), torch.zeros(B, cfg.n_heads, past_len, cfg.head_dim)),)
        ids = torch.full((B, 1), 5, dtype=torch.long)
        with torch.inference_mode():
            r = model.prepare_inputs_labels_for_multimodal(ids, None, mask, past, None, torch.zeros(B, 2, cfg.adapter_dim), None, None, None, None)
        rows[key + "_mask_in"], rows[key + "_past_len"] = mask.numpy(), np.int64(past_len)
        rows[key + "_mask_out"], rows[key + "_pos_out"] = r[2].numpy(), r[1].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rows)
    print(name, {k: v.tolist() for k, v in rows.items() if k.endswith("pos_out")})


def case_stage1_memory(ref, name, cfg, n_seg, n_frames, n_mem, n_prefix, steps):
    """The <memory> streaming branch (vtimellm_arch.py:208-232): ids hold -200 and -300, the memory block is
    [embed_tokens(prefix_memory) ; mm_projector(visual_memory)]."""
    w = syn.make_llama_weights(cfg, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w)
    feats = syn.make_features(n_seg, n_frames, cfg.adapter_dim, seed=21).float()
    vis_mem = syn.make_features(n_seg, n_mem, cfg.adapter_dim, seed=22).float()
    g = torch.Generator().manual_seed(23)
    prefix = torch.randint(3, cfg.vocab, (n_seg, n_prefix), generator=g)
    base = syn.make_prompt_ids(cfg, 6, 9, seed=24)
    base = torch.cat([base[:-4], torch.tensor([ref.constants.MEMORY_TOKEN_INDEX]), base[-4:]])     # ... <video> text <memory> text
    ids = base[None].repeat(n_seg, 1)
    with torch.inference_mode():
        r = model.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, feats, None, vis_mem, prefix, None)
        embeds = r[4]
        am = torch.ones(embeds.shape[:2], dtype=torch.bool)
        logits_all, toks, scores = greedy_via_reference(model, embeds, am, steps)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), digest=syn.weights_digest(w), feats=feats.numpy(), vis_mem=vis_mem.numpy(),
        prefix=prefix.numpy(), ids=ids.numpy(), embeds=embeds.numpy(), prefill_logits=logits_all.numpy().astype(np.float32),
        tokens=toks.numpy(), scores=scores.numpy())
    print(name, "embeds", tuple(embeds.shape), "tokens", toks[0].tolist())


def case_clip_encoder(ref, name, cfg, V, T, Lq):
    w = syn.make_llama_weights(cfg, seed=0)
    cw = syn.make_clip_encoder_weights(cfg.hidden, seed=0)
    model = ref_shim.build_reference_model(ref, cfg, w, clip_weights=cw)
    frames = syn.make_features(V, T, 768, seed=3).float()
    g = torch.Generator().manual_seed(5)
    q_tok = torch.randn(Lq, 768, generator=g).to(torch.bfloat16).float()
    q, qmask = ref.tensor_utils.pad_sequences_1d([q_tok, q_tok[: Lq - 2]], dtype=torch.float32)
    ids = syn.make_prompt_ids(cfg, 5, 8, seed=4)[None]
    with torch.inference_mode():
        enc = model.get_model().mm_projector
        # direct adapter call on V segments, ragged text mask (row 1 of the padded pair)
        txt = q[1:2].repeat(V, 1, 1)
        msk = qmask[1:2].repeat(V, 1)
        cls_out = enc(frames, txt, msk, None)                 # [V, 1, hidden]
        r = model.prepare_inputs_labels_for_multimodal(
            ids, None, None, None, None, frames[None], (q[0:1], qmask[0:1]), None, None, None)
        embeds = r[4]
        logits = model(inputs_embeds=embeds).logits[:, -1].float()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        digest=syn.weights_digest(w), clip_digest=syn.weights_digest(cw),
        frames=frames.numpy(), q=q.numpy(), qmask=qmask.numpy(), ids=ids.numpy(),
        cls_out=cls_out[:, 0].numpy(), embeds=embeds.numpy(), last_logits=logits.numpy(),
    )
    print(name, "cls_out", tuple(cls_out.shape), "embeds", tuple(embeds.shape))


def case_scoring(ref, name):
    g = torch.Generator().manual_seed(11)
    text = torch.randn(2, 768, generator=g)
    video = torch.randn(3, 9, 768, generator=g)
    pooled = ref.similarity._topk_pooling(text, video, 3)
    # stage-1 caller arithmetic (eval_nlq_negative.py:309-316) and stage-2 (e2e2:380-386)
    cls = torch.randn(768, generator=g)
    prop = torch.randn(17, 768, generator=g)
    pf0 = prop / prop.norm(dim=0, keepdim=True)
    s0 = torch.einsum("bd,d->b", ref.similarity._topk_pooling(cls[None], pf0[None], 3)[0], cls)
    pf1 = prop[None] / prop[None].norm(dim=1, keepdim=True)      # e2e2 quirk: dim=1 of [1,n,d] is the frame axis too
    s1 = torch.einsum("bd,d->b", ref.similarity._topk_pooling(cls[None], pf1, 3)[:, 0], cls)
    pf2 = prop / prop.norm(dim=1, keepdim=True)                  # similarity.py:61 per-frame normalisation
    s2 = torch.einsum("ld,d->l", ref.similarity._topk_pooling(cls[None], pf2[None], 3)[0], cls)
    logits = torch.randn(3, 5, 512, generator=g) * 3
    ent = ref.entropy.get_entropy_statistics(logits, 0, 5)
    ent1 = ref.entropy.get_entropy_statistics(logits[:, :1], 0, 1)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        text=text.numpy(), video=video.numpy(), pooled=pooled.numpy(),
        cls=cls.numpy(), prop=prop.numpy(), s_norm0=s0.numpy(), s_norm1_batched=s1.numpy(), s_perframe=s2.numpy(),
        logits=logits.numpy(), ent=ent.numpy(), ent1=ent1.float().numpy(),
    )
    print(name, "ok")


def case_prompt(ref, name):
    tok = syn.StubTokenizer(32000)
    rows = {}
    for key, q in (("stage1", "<video>\nDuring which frames can we see a man opens the door?"),
                   ("stage2", "<video>\nDuring which video can we see she picks up 2 cups?")):
        conv = ref.conversation.conv_templates["v1"].copy()
        conv.append_message(conv.roles[0], q)
        conv.append_message(conv.roles[1], None)
        prompt = conv.get_prompt()
        ids = ref.mm_utils.tokenizer_image_token(prompt, tok, ref.constants.IMAGE_TOKEN_INDEX, return_tensors="pt")
        rows[key + "_prompt"] = np.array(prompt)
        rows[key + "_ids"] = ids.numpy()
    # memory variant (inference.py:29-30)
    conv = ref.conversation.conv_templates["v1"].copy()
    conv.append_message(conv.roles[0], "<video>\nWhere is the cat?<memory>")
    conv.append_message(conv.roles[1], None)
    prompt = conv.get_prompt()
    rows["memory_prompt"] = np.array(prompt)
    rows["memory_ids"] = ref.mm_utils.tokenizer_image_token(prompt, tok, -200, return_tensors="pt").numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rows)
    print(name, {k: (v.shape if v.ndim else str(v)[:40]) for k, v in rows.items()})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = ref_shim.load()
    case_stage1(ref, "stage1_tiny", syn.TINY, n_seg=3, n_frames=20, n_pre=6, n_post=9, steps=6)
    case_stage1(ref, "stage1_ragged", syn.TINY, n_seg=4, n_frames=12, n_pre=5, n_post=11, steps=4, ragged=True)
    case_stage1_memory(ref, "stage1_memory", syn.TINY, n_seg=2, n_frames=10, n_mem=3, n_prefix=4, steps=4)
    if "--all" in sys.argv or "truncated" in sys.argv:
        case_stage1_truncated(ref, "stage1_truncated", syn.TINY, n_seg=4, n_frames=14, n_pre=5, n_post=12, max_len=17)
        case_stage1_truncated(ref, "stage1_truncated_text", syn.TINY, n_seg=3, n_frames=6, n_pre=5, n_post=12, max_len=19)
        case_stage1_no_placeholder(ref, "stage1_no_placeholder", syn.TINY, n_seg=4, n_frames=9, n_pre=5, n_post=8, rows_without=(1, 3))
        case_stage1_no_placeholder(ref, "stage1_no_placeholder_truncated", syn.TINY, n_seg=4, n_frames=9, n_pre=5, n_post=8, rows_without=(1, 3), max_len=11)
        case_stage1_image_list(ref, "stage1_image_list", syn.TINY, frames=(11, 3, 1, 17, 8), n_pre=5, n_post=9)
        case_decode_fixup(ref, "decode_fixup", syn.TINY)
        case_two_placeholders(ref, "stage1_two_placeholders", syn.TINY)
    case_clip_encoder(ref, "clip_encoder_tiny", syn.TINY, V=5, T=12, Lq=7)
    case_scoring(ref, "scoring")
    case_prompt(ref, "prompt")


if __name__ == "__main__":
    main()
