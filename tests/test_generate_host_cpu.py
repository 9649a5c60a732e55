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
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM


class CpuEngine:
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

    def ensure_kv(self, n_pages):
        self.n_pages = max(self.n_pages, n_pages)

    def project_splice(self, rows, vis_dst, text_ids, text_dst, hidden):
        hidden[text_dst.long()] = self.w["model.embed_tokens.weight"][text_ids.long()]
        proj = rows.float() @ self.w["model.mm_projector.weight"].t() + self.w["model.mm_projector.bias"]
        hidden[vis_dst.long()] = proj

    def _key(self, table_row, seq_len):
        # the page that holds the sequence's LAST prompt position is never a shared prefix page
        return int(table_row[(int(seq_len) - 1) // self.cfg.kv_page_size])

    def prefill(self, hidden, cu, n_seq, max_seqlen, page_table, logits_out, all_logits=False, seq_pos0=None, seq_ctx_row=None):
        assert seq_pos0 is None and not all_logits
        self.calls.append(("prefill", n_seq))
        for b in range(n_seq):
            x = hidden[int(cu[b]):int(cu[b + 1])][None]
            cache = llama_ref.KVCache(self.shape.n_layers)
            h = llama_ref.decoder_stack(self.w, self.shape, x, cache)
            logits_out[b] = llama_ref.lm_head(self.w, h[:, -1])[0]
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
