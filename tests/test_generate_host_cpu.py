"""generate()'s HOST logic on the CPU: `model.RevisionLlamaForCausalLM.generate` driven over a stand-in for the engine whose
entry points (project_splice, prefill, sample_greedy, decode_step, KV pool) are restated with the oracle's fp32 math.
What is under test is the Python around the C ABI - splice plan application, KV page tables, the step loop, EOS / pad
bookkeeping, retirement of finished rows, output layout - not the kernels (tests/test_gpu_*.py do that on the B200).
The stand-in lives in this test file; the product path never sees it (RevisionLlamaForCausalLM.cuda() builds the real
engine and fails loudly without an sm_100 device)."""
import contextlib
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import llama_ref, splice_ref
from revisionllm_b200 import synthetic as syn
from revisionllm_b200.engine import DecodeChunks
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM


class CpuEngine(DecodeChunks):
    device = torch.device("cpu")

    def __init__(self, cfg, w):
        self.w = {k: v.float() for k, v in w.items()}
        self.shape = llama_ref.LlamaShape(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate, cfg.vocab,
                                          cfg.rms_eps, cfg.rope_theta, cfg.adapter_dim)
        self.cfg = SimpleNamespace(kv_page_size=32, vocab=cfg.vocab, hidden=cfg.hidden)
        self.launches = 0
        self.caches = {}            # first OWN page id of a sequence -> its KV cache (pages are the sequence's identity)
        self.n_pages = 0
        self.calls = []
        self._init_chunks()

    def ensure_kv(self, n_pages):
        self.n_pages = max(self.n_pages, n_pages)

    def project_splice(self, rows, vis_dst, text_ids, text_dst, hidden):
        hidden[text_dst.long()] = self.w["model.embed_tokens.weight"][text_ids.long()]
        proj = rows.float() @ self.w["model.mm_projector.weight"].t() + self.w["model.mm_projector.bias"]
        hidden[vis_dst.long()] = proj

    def _key(self, table_row, seq_len):
        # the page that holds the sequence's LAST prompt position is never a shared prefix page
        return int(table_row[(int(seq_len) - 1) // self.cfg.kv_page_size])

    def splice_rows(self, vis, vis_dst, text_ids, text_dst, hidden):
        hidden[text_dst.long()] = self.w["model.embed_tokens.weight"][text_ids.long()]
        hidden[vis_dst.long()] = vis.float()

    def prefill(self, hidden, cu, n_seq, max_seqlen, page_table, logits_out, all_logits=False, seq_pos0=None, seq_ctx_row=None):
        self.calls.append(("prefill", n_seq))
        self.prefill_rows = int(cu[n_seq])
        self.caches.clear()                              # a prefill starts a new batch: page ids of the previous one mean nothing now
        for b in range(n_seq):
            x = hidden[int(cu[b]):int(cu[b + 1])][None]
            if seq_pos0 is not None and int(seq_pos0[b]) > 0:
                # positions [0, pos0) of this sequence are the rows [ctx_row, ctx_row + pos0) of the packed stream (include/
                # revisionllm_b200.h, rvl_prefill): the shared prompt prefix, embedded once
                c0, p0 = int(seq_ctx_row[b]), int(seq_pos0[b])
                x = torch.cat([hidden[c0:c0 + p0][None], x], dim=1)
            cache = llama_ref.KVCache(self.shape.n_layers)
            h = llama_ref.decoder_stack(self.w, self.shape, x, cache)
            if all_logits:
                logits_out[int(cu[b]):int(cu[b + 1])] = llama_ref.lm_head(self.w, h[0])
            else:
                logits_out[b] = llama_ref.lm_head(self.w, h[:, -1])[0]
            if seq_pos0 is not None and int(seq_pos0[b]) == 0:
                continue                                 # the shared prefix itself: its K/V live on in every sequence that names it
            self.caches[self._key(page_table[b], x.shape[1])] = (cache, x.shape[1])

    def decode_step(self, tok, seq_lens, page_table, logits, max_kv_len=0):
        self.calls.append(("decode", tok.shape[0]))
        for b in range(tok.shape[0]):
            cache, _ = self.caches[self._find(page_table[b])]
            assert int(seq_lens[b]) < page_table.shape[1] * self.cfg.kv_page_size, "page table too small for this position"
            h = llama_ref.decoder_stack(self.w, self.shape, llama_ref.embed_tokens(self.w, tok[b:b + 1].long())[:, None], cache)
            logits[b] = llama_ref.lm_head(self.w, h[:, -1])[0]
        seq_lens += 1                                    # the device increments the lengths at the end of a step

    def _find(self, table_row):
        for p in table_row.tolist():
            if p in self.caches:
                return p
        raise KeyError(table_row.tolist())

    # ---- the three entry points the ClipEncoder host code drives, with their documented meaning
    def gemm(self, A, W, bias=None, out=None, out_mode=0, flags=0, **kw):
        from revisionllm_b200 import _cabi
        y = A.float() @ W.float().t()
        if bias is not None:
            y = y + bias.float()
        if flags & _cabi.GEMM_FLAG_RELU:
            y = y.clamp(min=0)
        if out is None:
            out = torch.empty(y.shape, dtype=torch.bfloat16 if out_mode == _cabi.GEMM_OUT_BF16 else torch.float32)
        if out_mode == _cabi.GEMM_ADD_F32:
            out += y
        else:
            out.copy_(y.to(out.dtype))
        return out

    def layernorm(self, x, w=None, b=None, y_f32=None, y_bf16=None, pos=None, y_pos_bf16=None, period=0, eps=1e-5):
        y = x.float()
        if w is not None:
            y = torch.nn.functional.layer_norm(y, (x.shape[1],), w.float(), b.float(), eps)
        y = y.clone()
        if y_pos_bf16 is not None:
            y_pos_bf16.copy_((y + pos.float()[torch.arange(x.shape[0]) % period]).to(torch.bfloat16))
        if y_bf16 is not None:
            y_bf16.copy_(y.to(torch.bfloat16))
        if y_f32 is not None:
            y_f32.copy_(y)

    def mha96(self, q, k, v, out, n_seq, n_heads, Tq, Tk, kv_seq_idx=None, key_mask=None):
        d = q.shape[1] // n_heads
        qh = q.float().reshape(n_seq, Tq, n_heads, d).permute(0, 2, 1, 3)
        sel = torch.arange(n_seq) if kv_seq_idx is None else kv_seq_idx.long()
        kh = k.float().reshape(-1, Tk, n_heads, d)[sel].permute(0, 2, 1, 3)
        vh = v.float().reshape(-1, Tk, n_heads, d)[sel].permute(0, 2, 1, 3)
        s = qh @ kh.transpose(-1, -2) / float(d) ** 0.5
        if key_mask is not None:
            s = s.masked_fill(key_mask.reshape(-1, Tk)[sel][:, None, None, :] == 0, float("-inf"))
        o = torch.softmax(s, dim=-1) @ vh
        out.copy_(o.permute(0, 2, 1, 3).reshape(n_seq * Tq, n_heads * d).to(torch.bfloat16))

    def cosine_topk(self, frames, seg_offsets, cls, k=3, norm_axis=1, max_seg_rows=None, want_idx=True, seg_ends=None):
        from oracle import scoring_ref
        offs = seg_offsets.tolist()
        ends = offs[1:] if seg_ends is None else seg_ends.tolist()
        scores = [scoring_ref.cosine_topk_score(frames[offs[i]:ends[i]], cls, k, norm_axis=norm_axis)[0] for i in range(len(ends))]
        return torch.tensor(scores, dtype=torch.float32), None

    def select_topk(self, scores, k):
        from oracle import scoring_ref
        return torch.from_numpy(np.ascontiguousarray(scoring_ref.select_topk_segments(scores.numpy(), k))).to(torch.int32)

    def sample_multinomial(self, logits, next_tokens, temperature, seed, step, entropy=None, unfinished=None, eos_id=2, pad_id=2):
        from oracle import sampling_ref
        self.calls.append(("sample", (round(float(temperature), 6), int(seed), int(step))))
        toks, _ = sampling_ref.multinomial_draw(logits.numpy(), float(temperature), int(seed), int(step))
        nxt = torch.tensor(toks)
        if entropy is not None:
            p = torch.softmax(logits.float(), dim=-1)                 # the entropy is of the un-tempered distribution (raw scores, :321)
            entropy.copy_(-(p * torch.log(p + 1e-10)).sum(-1))
        if unfinished is not None and eos_id >= 0:
            nxt, unf = llama_ref.eos_bookkeeping(nxt, unfinished.long(), eos_id, pad_id)
            unfinished.copy_(unf.to(unfinished.dtype))
        next_tokens.copy_(nxt.to(next_tokens.dtype))

    def decode_n(self, n_steps, logits, token_ring, entropy_ring, unfinished, eos_id, pad_id, seq_lens, page_table, max_kv_len=0):
        for s in range(n_steps):                         # rvl_decode_n: the loop the C entry point runs
            self.sample_greedy(logits, token_ring[s], None if entropy_ring is None else entropy_ring[s], unfinished, eos_id, pad_id)
            self.decode_step(token_ring[s], seq_lens, page_table, logits, max_kv_len=max_kv_len)

    def sample_greedy(self, logits, next_tokens, entropy=None, unfinished=None, eos_id=2, pad_id=2):
        nxt = torch.argmax(logits, dim=-1)
        if entropy is not None:
            p = torch.softmax(logits.float(), dim=-1)
            entropy.copy_(-(p * torch.log(p + 1e-10)).sum(-1))
        if unfinished is not None and eos_id >= 0:
            tok, unf = llama_ref.eos_bookkeeping(nxt, unfinished.long(), eos_id, pad_id)
            unfinished.copy_(unf.to(unfinished.dtype))
            nxt = tok
        next_tokens.copy_(nxt.to(next_tokens.dtype))


@pytest.fixture()
def cpu_model(monkeypatch):
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    m = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), dict(w))
    m.engine = CpuEngine(cfg, w)
    m.device = torch.device("cpu")
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    return m, w, cfg


def _oracle(w, cfg, ids, feats, steps, **kw):
    shape = llama_ref.LlamaShape(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate, cfg.vocab, cfg.rms_eps,
                                 cfg.rope_theta, cfg.adapter_dim)
    x = torch.stack(splice_ref.splice(w, ids, splice_ref.mm_projector_linear(w, feats)))
    return llama_ref.greedy_decode(w, shape, x, steps, **kw)


def test_generate_loop_outputs_and_layout(cpu_model):
    m, w, cfg = cpu_model
    feats = syn.make_features(3, 9, cfg.adapter_dim, seed=1)
    ids = syn.make_prompt_ids(cfg, 6, 9, seed=2)[None].repeat(3, 1)
    out = m.generate(ids, images=feats, max_new_tokens=5, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    toks, scores = _oracle(w, cfg, ids, feats, 5, stop_on_eos=False)
    assert out["sequences"][:, : ids.shape[1]].tolist() == ids.tolist()                    # prompt ids echoed, -200 included
    assert out["sequences"][:, ids.shape[1]:].tolist() == toks.tolist()
    assert len(out["scores"]) == 5 and out["scores"][0].shape == (3, cfg.vocab)
    np.testing.assert_allclose(torch.stack(out["scores"]).numpy(), torch.stack(scores).numpy(), rtol=2e-4, atol=2e-4)
    assert out["entropies"].shape == (3, 5) and torch.isfinite(out["entropies"]).all()
    assert out["prompt_lengths"].tolist() == [ids.shape[1] - 1 + 9] * 3
    assert m.engine.calls == [("prefill", 3)] + [("decode", 3)] * 4                         # 5 tokens = prefill + 4 steps
    assert m.generate(ids, images=feats, max_new_tokens=2, eos_token_id=None).shape == (3, ids.shape[1] + 2)   # plain tensor otherwise


def test_generate_eos_padding_early_stop_and_retirement(cpu_model):
    m, w, cfg = cpu_model
    B = 4
    feats = syn.make_features(B, 7, cfg.adapter_dim, seed=5)
    ids = syn.make_prompt_ids(cfg, 6, 9, seed=6)[None].repeat(B, 1)
    free = m.generate(ids, images=feats, max_new_tokens=6, return_dict_in_generate=True, eos_token_id=None)["sequences"][:, ids.shape[1]:]
    eos = int(free[0, 2])                                  # row 0 "ends" at its third token; other rows may or may not emit it
    want, _ = _oracle(w, cfg, ids, feats, 6, eos_token_id=eos, pad_token_id=0, stop_on_eos=True)
    for retire in (False, True):
        m.engine.calls.clear()
        out = m.generate(ids, images=feats, max_new_tokens=6, output_scores=True, return_dict_in_generate=True, eos_token_id=eos,
                         pad_token_id=0, retire_finished=retire)
        got = out["sequences"][:, ids.shape[1]:]
        assert got.tolist() == want.tolist(), retire                                        # pad after EOS, stop when all finished
        assert len(out["scores"]) == want.shape[1]
        done_at = [(row == eos).nonzero()[0, 0].item() if (row == eos).any() else None for row in want]
        if retire and any(d is not None and d < want.shape[1] - 1 for d in done_at):
            sizes = [n for kind, n in m.engine.calls if kind == "decode"]
            assert sizes == sorted(sizes, reverse=True) and sizes[-1] < B                   # finished rows left the decode batch
            ent = out["entropies"]
            for b, d in enumerate(done_at):
                if d is not None and d + 1 < ent.shape[1]:
                    assert torch.isnan(ent[b, d + 1:]).all()                               # no entropies after a row retired


def test_generate_ragged_batch_and_argument_errors(cpu_model):
    from revisionllm_b200._cabi import RvlError
    m, w, cfg = cpu_model
    frames = [4, 11, 1]
    images = [syn.make_features(1, f, cfg.adapter_dim, seed=20 + i)[0] for i, f in enumerate(frames)]
    base = syn.make_prompt_ids(cfg, 5, 8, seed=7)
    ids = base[None].repeat(3, 1)
    am = torch.ones_like(ids, dtype=torch.bool)
    am[1, -2:] = False
    ids[1, -2:] = 0
    out = m.generate(ids, images=images, attention_mask=am, max_new_tokens=3, return_dict_in_generate=True, eos_token_id=None)
    shape = m.engine.shape
    proj = [splice_ref.mm_projector_linear(w, im[None].float())[0] for im in images]
    emb = splice_ref.splice(w, ids, proj, attention_mask=am)
    toks, _ = llama_ref.forward_ragged(w, shape, emb, 3, stop_on_eos=False)
    assert out["sequences"][:, ids.shape[1]:].tolist() == torch.stack(toks).tolist()
    assert out["prompt_lengths"].tolist() == [e.shape[0] for e in emb]
    with pytest.raises(RvlError):
        m.generate(ids, images=None)
    with pytest.raises(NotImplementedError):
        m.generate(ids, images=images, num_beams=2)
    with pytest.raises(RvlError):
        m.generate(ids, images=images, visual_memory=torch.zeros(3, 2, cfg.adapter_dim))


def test_forward_api_prefill_logits_and_one_token_steps(cpu_model):
    """forward(): all-position logits (right padded), logits_to_keep=1, then 1-token steps with past_key_values - the
    LlamaForCausalLM.forward surface the reference's generate() drives (vtimellm_llama.py:38-90)."""
    m, w, cfg = cpu_model
    feats = syn.make_features(2, 6, cfg.adapter_dim, seed=3)
    ids = syn.make_prompt_ids(cfg, 6, 9, seed=2)[None].repeat(2, 1)
    shape = m.engine.shape
    x = torch.stack(splice_ref.splice(w, ids, splice_ref.mm_projector_linear(w, feats)))
    cache = llama_ref.KVCache(shape.n_layers)
    ref_all = llama_ref.lm_head(w, llama_ref.decoder_stack(w, shape, x, cache))
    full = m.forward(input_ids=ids, images=feats)
    np.testing.assert_allclose(full.logits.numpy(), ref_all.numpy(), rtol=2e-4, atol=2e-4)
    last = m.forward(input_ids=ids, images=feats, logits_to_keep=1, reserve_new_tokens=8)
    np.testing.assert_allclose(last.logits[:, 0].numpy(), ref_all[:, -1].numpy(), rtol=2e-4, atol=2e-4)
    kv = last.past_key_values
    tok = ref_all[:, -1].argmax(-1)
    for _ in range(3):
        ref_step = llama_ref.lm_head(w, llama_ref.decoder_stack(w, shape, llama_ref.embed_tokens(w, tok)[:, None], cache)[:, -1])
        out = m.forward(input_ids=tok[:, None], past_key_values=kv)
        np.testing.assert_allclose(out.logits[:, 0].numpy(), ref_step.numpy(), rtol=2e-4, atol=2e-4)
        kv = out.past_key_values
        tok = ref_step.argmax(-1)
    assert kv.steps == 3 and kv.seq_lens.tolist() == [x.shape[1] + 3] * 2


def test_memory_branch_through_generate(cpu_model):
    """`visual_memory` / `prefix_memory` (vtimellm_arch.py:208-232) through generate(): the -300 placeholder becomes the
    prefix ids plus a second visual block; tokens equal the oracle's, which a reference fixture pins."""
    m, w, cfg = cpu_model
    B = 2
    feats = syn.make_features(B, 7, cfg.adapter_dim, seed=21)
    mem = syn.make_features(B, 3, cfg.adapter_dim, seed=22)
    prefix = torch.randint(3, cfg.vocab, (B, 4), generator=torch.Generator().manual_seed(23))
    base = syn.make_prompt_ids(cfg, 6, 9, seed=24)
    base = torch.cat([base[:-4], torch.tensor([-300]), base[-4:]])
    ids = base[None].repeat(B, 1)
    out = m.generate(ids, images=feats, visual_memory=mem, prefix_memory=prefix, max_new_tokens=3, return_dict_in_generate=True, eos_token_id=None)
    emb = splice_ref.splice(w, ids, splice_ref.mm_projector_linear(w, feats), visual_memory=mem.float(), prefix_memory=prefix)
    toks, _ = llama_ref.greedy_decode(w, m.engine.shape, torch.stack(emb), 3, stop_on_eos=False)
    assert out["sequences"][:, ids.shape[1]:].tolist() == toks.tolist()
    assert out["prompt_lengths"].tolist() == [e.shape[0] for e in emb]


def test_stage2_hierarchy_and_window_bank_through_generate(monkeypatch):
    """Stage-2 input through generate() on the CPU stand-ins: `images [b, v, t, 768]` + `query_feats` (one ClipEncoder CLS token
    per window spliced at <video>) equals the oracle's adapter + splice + greedy loop, and a `WindowBank` that stores each
    distinct (window, query) pair once - rows of different lengths included - gives the same tokens as the stacked copies."""
    from oracle import clip_encoder_ref
    from revisionllm_b200.clip_encoder import ClipEncoder
    from revisionllm_b200.model import WindowBank
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    cw = syn.make_clip_encoder_weights(cfg.hidden, seed=0)
    m = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg, clip_adapter=True, clip_adapter_text=True, hierarchy=True), dict(w), None)
    m.engine = CpuEngine(cfg, w)
    m.device = torch.device("cpu")
    m.clip_encoder = ClipEncoder(m.engine, cw)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    g = torch.Generator().manual_seed(3)
    wins = syn.make_features(5, 6, 768, seed=9)                                   # 5 distinct windows of 6 frames
    q_tok = torch.randn(2, 4, 768, generator=g).to(torch.bfloat16)
    q_mask = torch.ones(2, 4)
    q_mask[1, 3:] = 0
    ids = syn.make_prompt_ids(cfg, 6, 9, seed=2)
    rows = torch.tensor([[0, 0, 1, 1], [4, 2, 3, 3]])                             # zoom-style repeats; prompt 0 -> query 0, prompt 1 -> query 1
    stacked = wins[rows]                                                          # [2, 4, 6, 768]
    out_s = m.generate(ids[None].repeat(2, 1), images=stacked, query_feats=(q_tok, q_mask), max_new_tokens=3,
                       return_dict_in_generate=True, eos_token_id=None)
    # oracle: adapter per (window, its prompt's query) -> CLS rows -> splice -> greedy
    seg_text = torch.tensor([0, 0, 0, 0, 1, 1, 1, 1])
    cls = clip_encoder_ref.clip_encoder_cls(cw, stacked.reshape(8, 6, 768).float(), q_tok.float()[seg_text], q_mask[seg_text])
    x = torch.stack(splice_ref.splice(w, ids[None].repeat(2, 1), cls.reshape(2, 4, -1)))
    toks, _ = llama_ref.greedy_decode(w, m.engine.shape, x, 3, stop_on_eos=False)
    assert out_s["sequences"][:, ids.shape[0]:].tolist() == toks.tolist()
    # the bank: windows 0, 1 belong to query 0, windows 4, 2, 3 to query 1 - each pair once
    bank = WindowBank(windows=wins[torch.tensor([0, 1, 4, 2, 3])], rows=torch.tensor([[0, 0, 1, 1], [2, 3, 4, 4]]),
                      text_index=torch.tensor([0, 0, 1, 1, 1], dtype=torch.int32))
    out_b = m.generate(ids[None].repeat(2, 1), images=bank, query_feats=(q_tok, q_mask), max_new_tokens=3,
                       return_dict_in_generate=True, eos_token_id=None)
    assert out_b["sequences"].tolist() == out_s["sequences"].tolist()
    # prompts with different numbers of windows in one batch
    ragged = WindowBank(windows=bank.windows, rows=[torch.tensor([0, 0, 1, 1]), torch.tensor([2, 3])], text_index=bank.text_index)
    out_r = m.generate(ids[None].repeat(2, 1), images=ragged, query_feats=(q_tok, q_mask), max_new_tokens=3,
                       return_dict_in_generate=True, eos_token_id=None)
    assert out_r["prompt_lengths"].tolist() == [ids.shape[0] - 1 + 4, ids.shape[0] - 1 + 2]
    assert out_r["sequences"][0].tolist() == out_s["sequences"][0].tolist()
    alone = m.generate(ids[None], images=wins[torch.tensor([4, 2])][None], query_feats=(q_tok[1:2], q_mask[1:2]), max_new_tokens=3,
                       return_dict_in_generate=True, eos_token_id=None)
    assert out_r["sequences"][1].tolist() == alone["sequences"][0].tolist()


def test_stage2_pass_schedules_agree_on_the_cpu_model(monkeypatch):
    """sweep.stage2_pass over the real model class on the CPU stand-ins: all chunks of all zoom levels in one batched
    generate() with every distinct window through the adapter once == the reference's schedule (one generate() per chunk,
    stacked zoom repeats), a window count below one chunk included (the reference's negative-slice case)."""
    from revisionllm_b200 import sweep
    from revisionllm_b200.clip_encoder import ClipEncoder
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    m = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg, clip_adapter=True, clip_adapter_text=True, hierarchy=True), dict(w), None)
    m.engine = CpuEngine(cfg, w)
    m.device = torch.device("cpu")
    m.clip_encoder = ClipEncoder(m.engine, syn.make_clip_encoder_weights(cfg.hidden, seed=0))
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(1, 4, 768, generator=g).to(torch.bfloat16), torch.ones(1, 4))
    ids = syn.make_prompt_ids(cfg, 6, 9, seed=2)
    key = lambda res: [(r["zoom"], r["start"], r["perm"], r["tokens"], r["window"]) for r in res]
    for n_windows in (7, 3):
        wins = syn.make_features(n_windows, 5, 768, seed=30 + n_windows)
        kw = dict(grounding_windows=list(range(10, 10 + n_windows)), batch=4, zooms=(2, 1), max_new_tokens=3, perm_seed=1,
                  answer_number=lambda t: int(t[0]) % 4, eos_token_id=None)
        m.engine.calls.clear()
        fast = sweep.stage2_pass(m, wins, q, ids, **kw)
        n_fast = sum(1 for kind, _ in m.engine.calls if kind == "prefill")
        m.engine.calls.clear()
        slow = sweep.stage2_pass(m, wins, q, ids, max_calls_per_batch=1, dedup=False, **kw)
        n_slow = sum(1 for kind, _ in m.engine.calls if kind == "prefill")
        assert key(fast) == key(slow), n_windows
        assert n_fast == 1 and n_slow == len(slow)                              # one batched generate() against one per chunk
        assert all(r["window"] in kw["grounding_windows"] for r in fast)
    # with EOS on: a chunk's entropy statistics stop at its own EOS whatever else shares the batch (the reference's one
    # generate() per chunk, eval_nlq_retrieval_e2e2.py:353-359).  Three queries with different prompts answer differently; the
    # EOS id is the second token of the first query's answer, which the other two never emit.
    qs = []
    for k in range(3):
        n_windows = 5 + k
        qs.append(dict(windows=syn.make_features(n_windows, 5, 768, seed=60 + k), query_feats=q, input_ids=syn.make_prompt_ids(cfg, 6, 9, seed=20 + k),
                       grounding_windows=list(range(10, 10 + n_windows)), perm_seed=1))
    args = (4, (2, 1), 5, lambda t: int(t[0]) % 4)
    free = sweep.stage2_pass_queries(m, qs, *args, None)
    eos = int(free[0][0]["tokens"][1])
    assert all(eos not in r["tokens"] for res in free[1:] for r in res), "pick other prompt seeds"
    fast = sweep.stage2_pass_queries(m, qs, *args, eos)
    slow = sweep.stage2_pass_queries(m, qs, *args, eos, 1, dedup=False)
    assert all(r["tokens"][1] == eos for r in slow[0]) and all(eos not in r["tokens"] for res in slow[1:] for r in res)
    for qa, qb, qf in zip(fast, slow, free):
        for a, b, f in zip(qa, qb, qf):
            # (fp32 CPU matmuls are not batch-invariant to the last bit; a statistic over the wrong steps differs in the first digits)
            assert (a["inv_max_entropy"], a["inv_mean_entropy"]) == pytest.approx((b["inv_max_entropy"], b["inv_mean_entropy"]), rel=1e-4), (a, b)
            cut = b["tokens"].index(eos) + 1 if eos in b["tokens"] else len(b["tokens"])
            assert a["tokens"][:cut] == b["tokens"][:cut] and a["window"] == b["window"]
    # and the statistics of the stopped chunks are NOT those of the full five steps
    assert all(a["inv_mean_entropy"] != pytest.approx(f["inv_mean_entropy"], rel=1e-3) for a, f in zip(fast[0], free[0]))


def test_shared_prefix_compute_changes_nothing_but_the_rows_computed(cpu_model):
    """`share_prefix_compute`: the whole pages of the prompt prefix common to the batch are embedded once ([prefix | seq 0 from
    P | seq 1 from P | ...], B + 1 sequences with `seq_pos0` / `seq_ctx_row`).  On the CPU stand-in: same tokens and the same
    logits to the bit, (B - 1) * P fewer rows through the prefill, and the shared pages mapped once in the page table."""
    m, w, cfg = cpu_model
    B, F = 4, 6
    feats = syn.make_features(B, F, cfg.adapter_dim, seed=8)
    ids = syn.make_prompt_ids(cfg, 40, 7, seed=9)[None].repeat(B, 1)        # 40 text ids in front of <video>: P = 32
    kw = dict(images=feats, max_new_tokens=3, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    plain = m.generate(ids, **kw)
    rows_plain = m.engine.prefill_rows
    m.share_prefix_compute = True
    shared = m.generate(ids, **kw)
    rows_shared = m.engine.prefill_rows
    m.share_prefix_compute = False
    assert shared["sequences"].tolist() == plain["sequences"].tolist()
    assert torch.equal(torch.stack(shared["scores"]), torch.stack(plain["scores"]))
    assert rows_plain - rows_shared == (B - 1) * 32
    assert m.engine.calls.count(("prefill", B + 1)) == 1                    # the prefix went through as one more sequence
    table = shared["past_key_values"].page_table
    assert (table[:, 0] == table[0, 0]).all() and len(set(table[:, 1].tolist())) == B


def test_several_queries_on_the_same_segments_share_the_visual_context(cpu_model):
    """`image_index`: Q prompts (queries) per segment in one batch without repeating the features; with `share_prefix_compute`
    the system text AND the visual positions of a segment are embedded / run through the decoder once per segment (one context
    sequence per group of prompts).  Same tokens and logits - to the bit on the CPU stand-in - as the batch in which every
    prompt carries its own copy of the features, fewer rows through the prefill, the context pages mapped once per segment."""
    m, w, cfg = cpu_model
    S, Q, F, n_pre = 3, 2, 40, 30                                             # 30 text ids + 40 frames in front of the query: P = 64
    feats = syn.make_features(S, F, cfg.adapter_dim, seed=18)
    base = syn.make_prompt_ids(cfg, n_pre, 7, seed=19)
    g = torch.Generator().manual_seed(20)
    ids = []
    for s_ in range(S):                                                        # segment-major rows: (segment, query)
        for q_ in range(Q):
            row = base.clone()
            row[n_pre + 1:] = torch.randint(3, cfg.vocab, (row.shape[0] - n_pre - 1,), generator=torch.Generator().manual_seed(100 + q_))
            ids.append(row)
    ids = torch.stack(ids)
    index = torch.arange(S).repeat_interleave(Q)
    kw = dict(max_new_tokens=3, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
    plain = m.generate(ids, images=feats[index], **kw)                         # every prompt with its own copy of the features
    rows_plain = m.engine.prefill_rows
    indexed = m.generate(ids, images=feats, image_index=index, **kw)           # same batch, features stored once
    assert indexed["sequences"].tolist() == plain["sequences"].tolist()
    assert torch.equal(torch.stack(indexed["scores"]), torch.stack(plain["scores"]))
    m.share_prefix_compute = True
    shared = m.generate(ids, images=feats, image_index=index, **kw)
    rows_shared = m.engine.prefill_rows
    m.share_prefix_compute = False
    assert m.last_shared_prefix == 64
    assert shared["sequences"].tolist() == plain["sequences"].tolist()
    assert torch.equal(torch.stack(shared["scores"]), torch.stack(plain["scores"]))
    assert rows_plain - rows_shared == (S * Q - S) * 64                       # one context per segment instead of one per prompt
    assert m.engine.calls.count(("prefill", S + S * Q)) == 1
    table = shared["past_key_values"].page_table
    for s_ in range(S):                                                        # the two context pages: one copy per segment
        rows = table[s_ * Q:(s_ + 1) * Q, :2]
        assert (rows == rows[0]).all()
    assert len({tuple(r) for r in table[:, :2].tolist()}) == S and len(set(table[:, 2].tolist())) == S * Q
    # queries of different lengths (masked tails) in the same batch
    am = torch.ones_like(ids, dtype=torch.bool)
    am[1::2, -2:] = False
    plain_m = m.generate(ids, images=feats[index], attention_mask=am, **kw)
    m.share_prefix_compute = True
    shared_m = m.generate(ids, images=feats, image_index=index, attention_mask=am, **kw)
    m.share_prefix_compute = False
    assert torch.equal(torch.stack(shared_m["scores"]), torch.stack(plain_m["scores"]))
    assert shared_m["prompt_lengths"].tolist() == plain_m["prompt_lengths"].tolist()


def test_multi_query_sweep_equals_one_sweep_per_query(cpu_model):
    """sweep.score_segments_queries (Q queries over S segments in one pass, rows segment-major, features stored once, shared
    visual context) returns for every (segment, query) the record that sweep.score_segments returns for that query alone:
    same tokens, entropy statistics, spans and span-sliced cosine scores."""
    from revisionllm_b200 import sweep
    m, w, cfg = cpu_model
    S, Q, F = 5, 3, 40
    feats = syn.make_features(S, F, cfg.adapter_dim, seed=28)
    base = syn.make_prompt_ids(cfg, 30, 7, seed=29)
    ids = base[None].repeat(Q, 1)
    for q_ in range(1, Q):
        ids[q_, 31:] = torch.randint(3, cfg.vocab, (ids.shape[1] - 31,), generator=torch.Generator().manual_seed(200 + q_))
    cls = torch.randn(Q, cfg.adapter_dim, generator=torch.Generator().manual_seed(30)).to(torch.bfloat16)
    spans_of = lambda tok: torch.stack([tok[:, 0] % 7, tok[:, 0] % 7 + tok[:, 1] % 5], dim=1).to(torch.int32)    # a deterministic "parser"
    single = [sweep.score_segments(m, feats, ids[q_], cls[q_], 3, spans_of, None, None) for q_ in range(Q)]
    for share in (False, True):
        m.share_prefix_compute = share
        multi = sweep.score_segments_queries(m, feats, ids, cls, 3, spans_of, batch_segments=2, eos_token_id=None)
        m.share_prefix_compute = False
        assert multi.shape == (S * Q, sweep.REC_WORDS)
        for q_ in range(Q):
            a, b = sweep.unpack_records(multi[q_::Q]), sweep.unpack_records(single[q_])
            assert a["tokens"].tolist() == b["tokens"].tolist() and a["spans"].tolist() == b["spans"].tolist(), (share, q_)
            for k in ("h_mean", "h_max", "cos"):
                assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-6), (share, q_, k)
    assert m.last_shared_prefix == 64 or m.last_shared_prefix == 0


class PagedCpuEngine(CpuEngine):
    """Like CpuEngine, but K / V really live in the paged pool with the product's layout ([layer][k|v][page][head][slot][d],
    bf16, `engine._kv`) and every step reads them back THROUGH the page table - so a wrong table, a missing page after the pool
    grew, or a shared page that is not really shared shows up as wrong tokens."""

    def __init__(self, cfg, w):
        super().__init__(cfg, w)
        self._kv = None
        self.cfg.n_heads, self.cfg.head_dim, self.cfg.n_layers = cfg.n_heads, cfg.head_dim, cfg.n_layers

    def ensure_kv(self, n_pages):
        if self._kv is not None and self.n_pages >= n_pages:
            return
        c = self.cfg
        self._kv = torch.zeros(2 * c.n_layers * n_pages * c.n_heads * c.kv_page_size * c.head_dim, dtype=torch.bfloat16).view(torch.uint8)
        self.n_pages = n_pages

    def _pool(self, layer, which):
        c = self.cfg
        per = self.n_pages * c.n_heads * c.kv_page_size * c.head_dim
        flat = self._kv.view(torch.bfloat16)
        return flat[(2 * layer + which) * per:(2 * layer + which + 1) * per].view(self.n_pages, c.n_heads, c.kv_page_size, c.head_dim)

    def _store(self, table_row, cache, first_pos):
        ps = self.cfg.kv_page_size
        for layer in range(self.cfg.n_layers):
            for which, t in ((0, cache.k[layer]), (1, cache.v[layer])):
                pool = self._pool(layer, which)
                for p in range(first_pos, t.shape[2]):
                    pool[int(table_row[p // ps]), :, p % ps] = t[0, :, p].to(torch.bfloat16)

    def _load(self, table_row, n):
        ps = self.cfg.kv_page_size
        cache = llama_ref.KVCache(self.cfg.n_layers)
        pages = table_row[: (n + ps - 1) // ps].long()
        for layer in range(self.cfg.n_layers):
            k = self._pool(layer, 0)[pages].permute(1, 0, 2, 3).reshape(self.cfg.n_heads, -1, self.cfg.head_dim)[:, :n]
            v = self._pool(layer, 1)[pages].permute(1, 0, 2, 3).reshape(self.cfg.n_heads, -1, self.cfg.head_dim)[:, :n]
            cache.k[layer], cache.v[layer] = k[None].float(), v[None].float()
        return cache

    def prefill(self, hidden, cu, n_seq, max_seqlen, page_table, logits_out, all_logits=False, seq_pos0=None, seq_ctx_row=None):
        assert not all_logits
        self.calls.append(("prefill", n_seq))
        for b in range(n_seq):
            x = hidden[int(cu[b]):int(cu[b + 1])][None]
            p0 = int(seq_pos0[b]) if seq_pos0 is not None else 0
            if p0:                                       # positions [0, p0) are the context rows; only the own positions are stored
                c0 = int(seq_ctx_row[b])
                x = torch.cat([hidden[c0:c0 + p0][None], x], dim=1)
            cache = llama_ref.KVCache(self.shape.n_layers)
            h = llama_ref.decoder_stack(self.w, self.shape, x, cache)
            logits_out[b] = llama_ref.lm_head(self.w, h[:, -1])[0]
            self._store(page_table[b], cache, p0)

    def decode_step(self, tok, seq_lens, page_table, logits, max_kv_len=0):
        self.calls.append(("decode", tok.shape[0]))
        for b in range(tok.shape[0]):
            n = int(seq_lens[b])
            assert n < page_table.shape[1] * self.cfg.kv_page_size, "page table too small for this position"
            cache = self._load(page_table[b], n)
            h = llama_ref.decoder_stack(self.w, self.shape, llama_ref.embed_tokens(self.w, tok[b:b + 1].long())[:, None], cache)
            logits[b] = llama_ref.lm_head(self.w, h[:, -1])[0]
            self._store(page_table[b], cache, n)
        seq_lens += 1


def test_paged_kv_tables_shared_pages_and_growth_past_the_first_reservation(monkeypatch):
    """generate() over the paged stand-in: 40 common text ids in front of <video> put one whole page under every sequence
    (mapped once, `share_prefix_pages`), and 70 new tokens outgrow the first 64-token reservation so the pool is re-paged
    mid-generation (`_grow_kv`).  Tokens must equal the un-paged oracle loop in both cases, with and without page sharing."""
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    B = 3
    feats = syn.make_features(B, 5, cfg.adapter_dim, seed=12)
    ids = syn.make_prompt_ids(cfg, 40, 6, seed=13)[None].repeat(B, 1)
    for steps in (6, 70):
        want, _ = _oracle(w, cfg, ids, feats, steps, stop_on_eos=False)
        for share in (True, False):
            m = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), dict(w))
            m.engine = PagedCpuEngine(cfg, w)
            m.device = torch.device("cpu")
            m.share_prefix_pages = share
            out = m.generate(ids, images=feats, max_new_tokens=steps, return_dict_in_generate=True, eos_token_id=None)
            got = out["sequences"][:, ids.shape[1]:]
            same = (got == want).all(dim=0).long().cumprod(0).sum().item()          # leading steps on which every row agrees
            assert same == steps, (steps, share, same)                               # bf16 K/V storage must not flip a planted token
            table = out["past_key_values"].page_table
            assert (len(set(table[:, 0].tolist())) == 1) == share
            assert table.shape[1] * 32 >= ids.shape[1] - 1 + 5 + steps


def test_paged_kv_tables_with_one_shared_context_per_segment(monkeypatch):
    """Several prompts per segment over the paged stand-in (`image_index` + `share_prefix_compute`): the context pages of a
    segment are written once, by its context sequence, and read by all its prompts through the page table; 70 new tokens
    re-page the pool mid-generation with the groups intact.  Tokens equal the un-paged oracle loop on the repeated features."""
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    S, Q, F = 2, 3, 30
    feats = syn.make_features(S, F, cfg.adapter_dim, seed=22)
    base = syn.make_prompt_ids(cfg, 40, 6, seed=23)                               # 40 + 30 positions in front of the query: P = 64
    ids = base[None].repeat(S * Q, 1)
    for r in range(S * Q):
        if r % Q:
            ids[r, 41:-2] = torch.randint(3, cfg.vocab, (ids.shape[1] - 43,), generator=torch.Generator().manual_seed(700 + r % Q))
    index = torch.arange(S).repeat_interleave(Q)
    for steps in (5, 70):
        want, _ = _oracle(w, cfg, ids, feats[index], steps, stop_on_eos=False)
        m = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), dict(w))
        m.engine = PagedCpuEngine(cfg, w)
        m.device = torch.device("cpu")
        m.share_prefix_compute = True
        out = m.generate(ids, images=feats, image_index=index, max_new_tokens=steps, return_dict_in_generate=True, eos_token_id=None)
        assert m.last_shared_prefix == 64 and m.engine.calls[0] == ("prefill", S + S * Q)
        got = out["sequences"][:, ids.shape[1]:]
        same = (got == want).all(dim=0).long().cumprod(0).sum().item()
        assert same == steps, (steps, same)
        table = out["past_key_values"].page_table
        for s_ in range(S):
            blk = table[s_ * Q:(s_ + 1) * Q, :2]
            assert (blk == blk[0]).all()
        assert len({tuple(r) for r in table[:, :2].tolist()}) == S


def test_sampling_arguments_reach_the_engine(cpu_model):
    """`do_sample=True, temperature=0.05, seed=s` (the reference's own rule, inference.py:47-48): every step calls the
    multinomial entry point with (temperature, seed, step index) - the Philox counter layout documented in the header - and
    the draw is reproducible per seed; temperature 0 or do_sample=False fall back to greedy."""
    m, w, cfg = cpu_model
    feats = syn.make_features(3, 6, cfg.adapter_dim, seed=1)
    ids = syn.make_prompt_ids(cfg, 6, 9, seed=2)[None].repeat(3, 1)
    kw = dict(images=feats, max_new_tokens=4, return_dict_in_generate=True, eos_token_id=None)
    m.engine.calls.clear()
    a = m.generate(ids, do_sample=True, temperature=0.05, seed=11, **kw)
    assert [c[1] for c in m.engine.calls if c[0] == "sample"] == [(0.05, 11, t) for t in range(4)]
    b = m.generate(ids, do_sample=True, temperature=0.05, seed=11, **kw)
    assert a["sequences"].tolist() == b["sequences"].tolist()
    hot_a = m.generate(ids, do_sample=True, temperature=50.0, seed=1, **kw)["sequences"]
    hot_b = m.generate(ids, do_sample=True, temperature=50.0, seed=2, **kw)["sequences"]
    assert hot_a.tolist() != hot_b.tolist()                                  # a flat distribution: the seed decides
    greedy = m.generate(ids, **kw)["sequences"]
    m.engine.calls.clear()
    assert m.generate(ids, do_sample=True, temperature=0.0, **kw)["sequences"].tolist() == greedy.tolist()
    assert not [c for c in m.engine.calls if c[0] == "sample"]
    assert a["sequences"].tolist() == greedy.tolist()                        # at T = 0.05 the planted chain wins every draw


def test_stage1_sweep_records_over_the_cpu_model(cpu_model):
    """sweep.stage1_sweep over the real model class on the CPU stand-in: per-segment records (tokens, entropy statistics,
    cosine top-3 score) in batches, and the stage-2 selection on top - against the oracle's pieces."""
    from oracle import scoring_ref
    from revisionllm_b200 import sweep
    m, w, cfg = cpu_model
    n, F = 7, 6
    segs = syn.make_features(n, F, cfg.adapter_dim, seed=14)
    ids = syn.make_prompt_ids(cfg, 6, 9, seed=2)
    cls = torch.randn(cfg.adapter_dim, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    res = sweep.stage1_sweep(m, segs, ids, cls, max_new_tokens=4, batch=3, eos_token_id=None, stage2_topk=3)
    un = sweep.unpack_records(res.records)
    toks, scores = _oracle(w, cfg, ids[None].repeat(n, 1), segs, 4, stop_on_eos=False)
    assert un["tokens"][:, :4].tolist() == toks.tolist()
    ent = torch.stack([scoring_ref.step_entropy(s) for s in scores], dim=1)            # [n, steps]
    np.testing.assert_allclose(un["h_mean"].numpy(), ent.mean(1).numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(un["h_max"].numpy(), ent.max(1).values.numpy(), rtol=1e-4, atol=1e-5)
    # no span decoder: every window is scored over all its frames, normalised across the frame axis (eval_nlq_negative.py:311)
    want_cos = [scoring_ref.cosine_topk_score(segs[i], cls, 3, norm_axis=0)[0] for i in range(n)]
    np.testing.assert_allclose(un["cos"].numpy(), np.array(want_cos, dtype=np.float32), rtol=1e-6)
    assert res.stage2_indices.tolist() == scoring_ref.select_topk_segments(np.array(want_cos, dtype=np.float32), 3).tolist()
    assert [c for c in m.engine.calls if c[0] == "prefill"] == [("prefill", 3), ("prefill", 3), ("prefill", 1)]
    # with predicted spans the cosine score is taken over `features[k][from : to + 1]` (eval_nlq_negative.py:309-310): a
    # one-frame span widened by one frame (:90-92), slices clamped at the window's end, windows without a span scored whole
    table = torch.tensor([[1, 3], [2, 2], [0, 0], [-1, -1], [4, 9], [5, 5], [0, 5]], dtype=torch.int32)
    rows_seen = []

    def decode_spans(tok):
        rows_seen.append(tok.shape[0])
        start = sum(rows_seen[:-1])
        return table[start: start + tok.shape[0]]
    res2 = sweep.stage1_sweep(m, segs, ids, cls, max_new_tokens=4, batch=3, eos_token_id=None, decode_spans=decode_spans)
    un2 = sweep.unpack_records(res2.records)
    assert un2["spans"].tolist() == table.tolist()
    slices = [(1, 4), (1, 4), (0, 2), (0, F), (4, F), (4, F), (0, F)]
    want2 = [scoring_ref.cosine_topk_score(segs[i][a:b], cls, 3, norm_axis=0)[0] for i, (a, b) in enumerate(slices)]
    np.testing.assert_allclose(un2["cos"].numpy(), np.array(want2, dtype=np.float32), rtol=1e-6)
    # a rank with an empty shard (fewer windows than ranks) contributes zero records and still returns
    assert sweep.score_segments(m, segs[:0], ids, cls, 4, eos_token_id=None).shape == (0, sweep.REC_WORDS)
