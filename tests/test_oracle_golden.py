"""Pin the CPU oracle (oracle/) against vectors produced by the reference itself
(tests/golden/make_golden.py ran the reference's own modules from /root/reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_encoder_ref, llama_ref, scoring_ref, splice_ref
from revisionllm_b200 import synthetic as syn


def _shape(cfg):
    return llama_ref.LlamaShape(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate,
                                cfg.vocab, cfg.rms_eps, cfg.rope_theta, cfg.adapter_dim)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=True)


def test_weights_generator_is_stable(golden_dir):
    g = _load(golden_dir, "stage1_tiny")
    w = syn.make_llama_weights(syn.TINY, seed=0)
    assert syn.weights_digest(w) == str(g["digest"])


def test_stage1_splice_prefill_and_greedy_match_reference(golden_dir):
    g = _load(golden_dir, "stage1_tiny")
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    feats, ids = torch.from_numpy(g["feats"]), torch.from_numpy(g["ids"])
    img = splice_ref.mm_projector_linear(w, feats)
    emb = splice_ref.splice(w, ids, img)
    x = torch.stack(emb)
    np.testing.assert_allclose(x.numpy(), g["embeds"], rtol=1e-5, atol=1e-5)
    hidden = llama_ref.decoder_stack(w, _shape(cfg), x)
    logits = llama_ref.lm_head(w, hidden)
    np.testing.assert_allclose(logits.numpy(), g["prefill_logits"], rtol=2e-4, atol=2e-4)
    steps = g["tokens"].shape[1]
    toks, scores = llama_ref.greedy_decode(w, _shape(cfg), x, steps, stop_on_eos=False)
    assert toks.tolist() == g["tokens"].tolist()
    np.testing.assert_allclose(torch.stack(scores).numpy(), g["scores"], rtol=2e-4, atol=2e-4)
    # planted successor chain (size-independent property used at full size on the GPU)
    succ = syn.successor_table(cfg)
    cur = int(ids[0, -1])
    for t in range(steps):
        cur = int(succ[cur])
        assert int(toks[0, t]) == cur


def test_ragged_batch_matches_reference_right_padding(golden_dir):
    g = _load(golden_dir, "stage1_ragged")
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    feats, ids, attn = torch.from_numpy(g["feats"]), torch.from_numpy(g["ids"]), torch.from_numpy(g["attn"])
    emb = splice_ref.splice(w, ids, splice_ref.mm_projector_linear(w, feats), attention_mask=attn)
    x, m, p = splice_ref.right_pad(emb)
    np.testing.assert_allclose(x.numpy(), g["embeds"], rtol=1e-5, atol=1e-5)
    assert m.numpy().tolist() == g["embeds_mask"].tolist()
    steps = g["tokens"].shape[1]
    toks, scores = llama_ref.forward_ragged(w, _shape(cfg), emb, steps, stop_on_eos=False)
    assert torch.stack(toks).tolist() == g["tokens"].tolist()
    sc = torch.stack([torch.stack(s) for s in scores], dim=1)      # [T, B, V]
    np.testing.assert_allclose(sc.numpy(), g["scores"], rtol=2e-4, atol=2e-4)
    am, pos = splice_ref.decode_step_fixup(m.long(), x.shape[1])
    assert pos[:, 0].tolist() == [e.shape[0] for e in emb]


def test_clip_encoder_and_hierarchy_splice_match_reference(golden_dir):
    g = _load(golden_dir, "clip_encoder_tiny")
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    cw = syn.make_clip_encoder_weights(cfg.hidden, seed=0)
    assert syn.weights_digest(cw) == str(g["clip_digest"])
    frames, q, qmask = torch.from_numpy(g["frames"]), torch.from_numpy(g["q"]), torch.from_numpy(g["qmask"])
    V = frames.shape[0]
    out = clip_encoder_ref.clip_encoder_cls(cw, frames, q[1:2].repeat(V, 1, 1), qmask[1:2].repeat(V, 1))
    np.testing.assert_allclose(out.numpy(), g["cls_out"], rtol=2e-4, atol=2e-4)
    feats = clip_encoder_ref.hierarchy_features(cw, frames[None], (q[0:1], qmask[0:1]))
    emb = splice_ref.splice(w, torch.from_numpy(g["ids"]), feats)
    np.testing.assert_allclose(torch.stack(emb).numpy(), g["embeds"], rtol=2e-4, atol=2e-4)
    hidden = llama_ref.decoder_stack(w, _shape(cfg), torch.stack(emb))
    np.testing.assert_allclose(llama_ref.lm_head(w, hidden[:, -1]).numpy(), g["last_logits"], rtol=5e-4, atol=5e-4)


def test_topk_pooling_and_entropy_match_reference(golden_dir):
    g = _load(golden_dir, "scoring")
    pooled = scoring_ref.topk_pooling(torch.from_numpy(g["text"]), torch.from_numpy(g["video"]), 3)
    np.testing.assert_allclose(pooled.numpy(), g["pooled"], rtol=1e-5, atol=1e-5)
    cls, prop = torch.from_numpy(g["cls"]), torch.from_numpy(g["prop"])
    s0, _, _ = scoring_ref.cosine_topk_score(prop, cls, 3, norm_axis=0)
    np.testing.assert_allclose(s0, g["s_norm0"][0], rtol=1e-5)
    np.testing.assert_allclose(s0, g["s_norm1_batched"][0], rtol=1e-5)   # e2e2's dim=1 on [1,n,d] is the same axis
    s2, _, _ = scoring_ref.cosine_topk_score(prop, cls, 3, norm_axis=1)
    np.testing.assert_allclose(s2, g["s_perframe"][0], rtol=1e-5)
    ent = scoring_ref.get_entropy_statistics(torch.from_numpy(g["logits"]))
    np.testing.assert_allclose(ent.numpy(), g["ent"], rtol=1e-5, atol=1e-6)
    ent1 = scoring_ref.get_entropy_statistics(torch.from_numpy(g["logits"])[:, :1])
    np.testing.assert_allclose(ent1.numpy(), g["ent1"], rtol=1e-5, atol=1e-6)


def test_prompt_and_placeholder_tokenisation_match_reference(golden_dir):
    g = _load(golden_dir, "prompt")
    tok = syn.StubTokenizer(32000)
    for key, q in (("stage1", "<video>\nDuring which frames can we see a man opens the door?"),
                   ("stage2", "<video>\nDuring which video can we see she picks up 2 cups?")):
        prompt = splice_ref.vicuna_v1_prompt(q)
        assert prompt == str(g[key + "_prompt"])
        assert splice_ref.tokenizer_image_token(prompt, tok) == g[key + "_ids"].tolist()
    prompt = splice_ref.vicuna_v1_prompt("<video>\nWhere is the cat?<memory>")
    assert prompt == str(g["memory_prompt"])
    assert splice_ref.tokenizer_image_token(prompt, tok) == g["memory_ids"].tolist()


def test_windows_and_selection_rules():
    # eval_nlq_negative.py:224-235 on a 1 h MAD movie: 18000 feats, 625-feat windows -> 57 windows
    w = scoring_ref.stage1_windows(18000, 625, 250)
    assert w.shape == (57, 250)
    assert w[1, 0] == 312 and w[0, -1] == 625
    w2, times = scoring_ref.stage2_windows(18000, 625, 250, stride=5)
    assert w2.shape == (143, 250)
    assert all(e - s == 625 for s, e in times)
    assert scoring_ref.nonoverlap_segments(18000, 100).shape == (180, 100)
    sel = scoring_ref.stage2_select_windows(["Not Present", "From 3 to 9.", "Not Present", "From 1 to 2."], 40, 10)
    assert len(sel) == 10 and sel == sorted(sel)
    assert scoring_ref.parse_span("From 12 to 34.") == (12, 34)
    assert scoring_ref.parse_span("Not Present") is None
    assert scoring_ref.select_topk_segments(np.array([1.0, 3.0, 3.0, 2.0], np.float32), 2).tolist() == [1, 2]


def test_iou_merge_and_ranking_metrics_match_reference(golden_dir):
    """oracle/metrics_ref.py against tests/golden/merge_metrics.json: the reference's own `iou` outputs and the metrics its
    metric_retrieval_forward.py script wrote for the same synthetic prediction files."""
    import json
    from oracle import metrics_ref
    g = json.load(open(os.path.join(golden_dir, "merge_metrics.json")))
    for c in g["iou_cases"]:
        cf, ious, kept = metrics_ref.iou(c["answers"], tuple(c["gt"]), c["num_frames_clip"], c["num_frames_video"], c["scores"],
                                         c["plus_baseline"])
        assert {str(k): list(v) for k, v in cf.items()} == c["clip_frames"]
        assert ious == c["ious"] and kept == c["kept_scores"]
    merged = [metrics_ref.merge_with_retrieval(q["gl"], q["rl"], q["rl2"]) for q in g["queries"]]
    got = metrics_ref.grounding_metrics_stream(merged)
    assert set(got) == set(g["metrics"])
    for k, v in g["metrics"].items():
        assert abs(got[k] - v) < 1e-9, (k, got[k], v)
    ratio = sum(len(m["answer"]) for m in merged) / sum(len(q["gl"]["answer"]) for q in g["queries"])
    assert abs(ratio - g["selected_ratio"]) < 1e-12


def test_memory_branch_splice_matches_reference(golden_dir):
    """<memory> streaming branch of the splice (vtimellm_arch.py:208-232) against the fixture recorded from the reference."""
    g = np.load(os.path.join(golden_dir, "stage1_memory.npz"))
    w = syn.make_llama_weights(syn.TINY, seed=0)
    assert syn.weights_digest(w) == str(g["digest"])
    ids, feats = torch.from_numpy(g["ids"]), torch.from_numpy(g["feats"])
    rows = splice_ref.splice(w, ids, splice_ref.mm_projector_linear(w, feats), visual_memory=torch.from_numpy(g["vis_mem"]),
                             prefix_memory=torch.from_numpy(g["prefix"]))
    got = torch.stack(rows)
    assert got.shape == tuple(g["embeds"].shape)
    np.testing.assert_allclose(got.numpy(), g["embeds"], rtol=1e-5, atol=1e-5)
    shape = llama_ref.LlamaShape(syn.TINY.hidden, syn.TINY.n_layers, syn.TINY.n_heads, syn.TINY.head_dim, syn.TINY.intermediate,
                                 syn.TINY.vocab, syn.TINY.rms_eps, syn.TINY.rope_theta, syn.TINY.adapter_dim)
    toks, scores = llama_ref.greedy_decode(w, shape, got, g["tokens"].shape[1], stop_on_eos=False)
    assert toks.tolist() == g["tokens"].tolist()
    np.testing.assert_allclose(torch.stack(scores).numpy(), g["scores"], rtol=2e-3, atol=2e-3)


def test_philox_known_answers_and_sampling_rule():
    from oracle import sampling_ref
    for ctr, key, out in sampling_ref.KNOWN_ANSWERS:
        assert sampling_ref.philox4x32_10(ctr, key) == out
    # a peaked distribution at T = 0.05 is the argmax; a flat one follows the uniform variate
    logits = np.zeros((4, 16), dtype=np.float32)
    logits[:, 5] = 3.0
    toks, _ = sampling_ref.multinomial_draw(logits, 0.05, seed=1, step=0)
    assert toks == [5, 5, 5, 5]
    flat = np.zeros((64, 8), dtype=np.float32)
    toks, _ = sampling_ref.multinomial_draw(flat, 1.0, seed=3, step=2)
    assert toks == [int(sampling_ref.uniform(3, 2, b) * 8) for b in range(64)]


@pytest.mark.parametrize("name", ["stage1_truncated", "stage1_truncated_text"])
def test_truncation_to_tokenizer_model_max_length_matches_reference(golden_dir, name):
    """`config.tokenizer_model_max_length` (vtimellm_arch.py:239-243) cuts every spliced row - inside the visual block in the
    first fixture, inside the trailing text / not at all in the second.  The oracle's splice and the product's host index
    plan (`engine.plan_splice`, replayed here with plain gathers) must both reproduce the reference's padded embeddings."""
    from revisionllm_b200.engine import plan_splice
    g = _load(golden_dir, name)
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    assert syn.weights_digest(w) == str(g["digest"])
    feats, ids, attn = torch.from_numpy(g["feats"]), torch.from_numpy(g["ids"]), torch.from_numpy(g["attn"])
    max_len = int(g["max_len"])
    img = splice_ref.mm_projector_linear(w, feats)
    emb = splice_ref.splice(w, ids, img, attention_mask=attn, max_length=max_len)
    x, m, _ = splice_ref.right_pad(emb)
    np.testing.assert_allclose(x.numpy(), g["embeds"], rtol=1e-5, atol=1e-5)
    assert m.numpy().tolist() == g["embeds_mask"].tolist()
    # the product's plan: text ids -> embedding rows, visual source rows -> projected rows, both scattered to packed rows
    plan = plan_splice(ids.numpy(), [feats.shape[1]] * feats.shape[0], attn.numpy(), max_length=max_len)
    assert plan["lengths"].tolist() == g["embeds_mask"].sum(1).tolist()
    T = int(plan["cu_seqlens"][-1])
    packed = torch.zeros(T, x.shape[2])
    packed[torch.from_numpy(plan["text_dst"]).long()] = w["model.embed_tokens.weight"].float()[torch.from_numpy(plan["text_ids"]).long()]
    packed[torch.from_numpy(plan["vis_dst"]).long()] = img.reshape(-1, img.shape[-1]).float()[torch.from_numpy(plan["vis_src"]).long()]
    cu = plan["cu_seqlens"]
    for b in range(ids.shape[0]):
        np.testing.assert_allclose(packed[cu[b]:cu[b + 1]].numpy(), g["embeds"][b, : cu[b + 1] - cu[b]], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["stage1_no_placeholder", "stage1_no_placeholder_truncated"])
def test_rows_without_placeholder_match_reference(golden_dir, name):
    """vtimellm_arch.py:168-176: a row without <video> is text only and still consumes its visual block, so the next rows
    keep theirs.  Oracle splice and the product's index plan against the reference's padded embeddings."""
    from revisionllm_b200.engine import plan_splice
    g = _load(golden_dir, name)
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    assert syn.weights_digest(w) == str(g["digest"])
    feats, ids = torch.from_numpy(g["feats"]), torch.from_numpy(g["ids"])
    max_len = None if int(g["max_len"]) < 0 else int(g["max_len"])        # the truncated fixture cuts the text-only rows too (:239-243)
    img = splice_ref.mm_projector_linear(w, feats)
    emb = splice_ref.splice(w, ids, img, max_length=max_len)
    lens = [e.shape[0] for e in emb]
    full = [ids.shape[1], ids.shape[1] - 1 + feats.shape[1]]                                  # text-only rows are shorter
    assert sorted(set(lens)) == sorted({min(n, max_len) if max_len else n for n in full})
    x, _, _ = splice_ref.right_pad(emb)
    np.testing.assert_allclose(x.numpy(), g["embeds"], rtol=1e-5, atol=1e-5)                  # zero padding behind the short rows
    plan = plan_splice(ids.numpy(), [feats.shape[1]] * feats.shape[0], max_length=max_len)
    assert plan["lengths"].tolist() == lens
    packed = torch.zeros(int(plan["cu_seqlens"][-1]), x.shape[2])
    packed[torch.from_numpy(plan["text_dst"]).long()] = w["model.embed_tokens.weight"].float()[torch.from_numpy(plan["text_ids"]).long()]
    packed[torch.from_numpy(plan["vis_dst"]).long()] = img.reshape(-1, img.shape[-1]).float()[torch.from_numpy(plan["vis_src"]).long()]
    cu = plan["cu_seqlens"]
    for b in range(ids.shape[0]):
        np.testing.assert_allclose(packed[cu[b]:cu[b + 1]].numpy(), g["embeds"][b, : lens[b]], rtol=1e-5, atol=1e-5)


def test_list_of_images_with_ragged_frame_counts_matches_reference(golden_dir):
    """vtimellm_arch.py:102-109: `images` as a list of [F_i, 768] tensors (one projector call, split per row) with right-padded
    prompts - the VidChapters-shaped batch.  Oracle splice and the product's index plan against the reference's embeddings."""
    from revisionllm_b200.engine import plan_splice
    g = _load(golden_dir, "stage1_image_list")
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    assert syn.weights_digest(w) == str(g["digest"])
    frames = [int(f) for f in g["frames"]]
    flat = torch.from_numpy(g["images"])
    images = list(torch.split(flat, frames, dim=0))
    ids, attn = torch.from_numpy(g["ids"]), torch.from_numpy(g["attn"])
    proj = [splice_ref.mm_projector_linear(w, im[None])[0] for im in images]
    emb = splice_ref.splice(w, ids, proj, attention_mask=attn)
    x, m, _ = splice_ref.right_pad(emb)
    np.testing.assert_allclose(x.numpy(), g["embeds"], rtol=1e-5, atol=1e-5)
    assert m.numpy().tolist() == g["embeds_mask"].tolist()
    plan = plan_splice(ids.numpy(), frames, attn.numpy())
    assert plan["lengths"].tolist() == g["embeds_mask"].sum(1).tolist()
    allproj = torch.cat(proj).float()
    packed = torch.zeros(int(plan["cu_seqlens"][-1]), x.shape[2])
    packed[torch.from_numpy(plan["text_dst"]).long()] = w["model.embed_tokens.weight"].float()[torch.from_numpy(plan["text_ids"]).long()]
    packed[torch.from_numpy(plan["vis_dst"]).long()] = allproj[torch.from_numpy(plan["vis_src"]).long()]
    cu = plan["cu_seqlens"]
    for b in range(ids.shape[0]):
        np.testing.assert_allclose(packed[cu[b]:cu[b + 1]].numpy(), g["embeds"][b, : cu[b + 1] - cu[b]], rtol=1e-5, atol=1e-5)


def test_decode_step_fixup_matches_reference(golden_dir):
    """vtimellm_arch.py:88-100 run by the reference itself: mask extended to cache length + 1, position = sum(mask) - 1.
    For right-padded rows that is the row's OWN length plus the steps already decoded - the `seq_lens` the CUDA decode step
    uses as position and KV slot of the new token."""
    g = _load(golden_dir, "decode_fixup")
    for key, lens, done in (("step1", (9, 6, 9, 4), 0), ("step3", (9, 6, 9, 4), 2), ("full", (7, 7), 0)):
        am, pos = splice_ref.decode_step_fixup(torch.from_numpy(g[key + "_mask_in"]), int(g[key + "_past_len"]))
        assert am.numpy().tolist() == g[key + "_mask_out"].tolist()
        assert pos.numpy().tolist() == g[key + "_pos_out"].tolist()
        assert [[n + done] for n in lens] == g[key + "_pos_out"].tolist()


def test_eos_bookkeeping_matches_reference_lines(golden_dir):
    """llama_ref.eos_bookkeeping (the rule inside the oracle's greedy loop, and the one `sample_greedy_kernel` implements on the
    device) against tests/golden/eos_rule.json: the reference's own `sample()` lines 340-362, exec'd step by step."""
    import json
    g = json.load(open(os.path.join(golden_dir, "eos_rule.json")))
    assert any(c["stopped_after_step"] is not None for c in g["cases"]) and any(c["stopped_after_step"] is None for c in g["cases"])
    for c in g["cases"]:
        raw = torch.tensor(c["raw"])
        unfinished = torch.ones(raw.shape[1], dtype=torch.long)
        for t in range(len(c["appended"])):
            tok, unfinished = llama_ref.eos_bookkeeping(raw[t], unfinished, c["eos"], c["pad"])
            assert tok.tolist() == c["appended"][t] and unfinished.tolist() == c["unfinished"][t], (c["eos"], t)
            if int(unfinished.max()) == 0:
                assert c["stopped_after_step"] == t
                break
        else:
            assert c["stopped_after_step"] is None


def test_two_placeholders_in_one_row_match_reference(golden_dir):
    """vtimellm_arch.py:178-207: a row with two <video> placeholders consumes two visual blocks and the next row the block
    after them.  Oracle splice and the product's index plan (general path) against the reference's padded embeddings."""
    from revisionllm_b200.engine import plan_splice
    g = _load(golden_dir, "stage1_two_placeholders")
    w = syn.make_llama_weights(syn.TINY, seed=0)
    assert syn.weights_digest(w) == str(g["digest"])
    frames = [int(f) for f in g["frames"]]
    images = list(torch.split(torch.from_numpy(g["images"]), frames, dim=0))
    ids = torch.from_numpy(g["ids"])
    proj = [splice_ref.mm_projector_linear(w, im[None])[0] for im in images]
    emb = splice_ref.splice(w, ids, proj)
    assert [e.shape[0] for e in emb] == [ids.shape[1] - 2 + frames[0] + frames[1], ids.shape[1] - 1 + frames[2]]
    x, _, _ = splice_ref.right_pad(emb)
    np.testing.assert_allclose(x.numpy(), g["embeds"], rtol=1e-5, atol=1e-5)
    plan = plan_splice(ids.numpy(), frames)
    assert plan["lengths"].tolist() == [e.shape[0] for e in emb]
    allproj = torch.cat(proj).float()
    packed = torch.zeros(int(plan["cu_seqlens"][-1]), x.shape[2])
    packed[torch.from_numpy(plan["text_dst"]).long()] = w["model.embed_tokens.weight"].float()[torch.from_numpy(plan["text_ids"]).long()]
    packed[torch.from_numpy(plan["vis_dst"]).long()] = allproj[torch.from_numpy(plan["vis_src"]).long()]
    cu = plan["cu_seqlens"]
    for b in range(2):
        np.testing.assert_allclose(packed[cu[b]:cu[b + 1]].numpy(), g["embeds"][b, : cu[b + 1] - cu[b]], rtol=1e-5, atol=1e-5)
