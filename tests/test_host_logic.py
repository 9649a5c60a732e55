"""CPU tests of the host-side mirror (no GPU): prompt/tokenisation, splice planning, windows,
selection rules, record packing, C-ABI surface."""
import ctypes
import math
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from oracle import scoring_ref, splice_ref
from revisionllm_b200 import _cabi, conversation, mm_utils, scoring, sweep, synthetic as syn
from revisionllm_b200.engine import plan_splice

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prompt_and_tokenizer_match_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "prompt.npz"), allow_pickle=True)
    tok = syn.StubTokenizer(32000)
    for key, q in (("stage1", "<video>\nDuring which frames can we see a man opens the door?"),
                   ("stage2", "<video>\nDuring which video can we see she picks up 2 cups?"),
                   ("memory", "<video>\nWhere is the cat?<memory>")):
        conv = conversation.conv_templates["v1"].copy()
        conv.append_message(conv.roles[0], q)
        conv.append_message(conv.roles[1], None)
        prompt = conv.get_prompt()
        assert prompt == str(g[key + "_prompt"])
        ids = mm_utils.tokenizer_image_token(prompt, tok, -200, return_tensors="pt")
        assert ids.tolist() == g[key + "_ids"].tolist()
    # the template object must not leak messages between copies
    assert conversation.conv_templates["v1"].messages == []


@pytest.mark.parametrize("ragged", [False, True])
def test_plan_splice_matches_oracle_splice(ragged):
    cfg = syn.TINY
    w = syn.make_llama_weights(cfg, seed=0)
    B, F = 4, 7
    ids = syn.make_prompt_ids(cfg, 5, 9, seed=3)[None].repeat(B, 1)
    attn = None
    if ragged:
        attn = torch.ones_like(ids, dtype=torch.bool)
        for b in range(B):
            cut = (2 * b) % 5
            if cut:
                attn[b, -cut:] = False
    feats = syn.make_features(B, F, cfg.adapter_dim, seed=4)
    img = splice_ref.mm_projector_linear(w, feats)
    ref = splice_ref.splice(w, ids, img, attention_mask=attn, max_length=18 if ragged else None)
    plan = plan_splice(ids.numpy(), [F] * B, None if attn is None else attn.numpy(), 18 if ragged else None)
    assert plan["lengths"].tolist() == [e.shape[0] for e in ref]
    # rebuild the packed stream from the plan with the oracle's embeddings and compare
    T = int(plan["cu_seqlens"][-1])
    packed = torch.zeros(T, cfg.hidden)
    emb = w["model.embed_tokens.weight"].float()
    packed[torch.from_numpy(plan["text_dst"]).long()] = emb[torch.from_numpy(plan["text_ids"]).long()]
    packed[torch.from_numpy(plan["vis_dst"]).long()] = img.reshape(B * F, -1)[torch.from_numpy(plan["vis_src"]).long()]
    np.testing.assert_array_equal(packed.numpy(), torch.cat(ref).numpy())


def test_plan_splice_row_without_placeholder_consumes_a_block():
    ids = np.array([[1, 5, 6, 7], [1, -200, 8, 9]], dtype=np.int64)
    plan = plan_splice(ids, [3, 3])
    assert plan["lengths"].tolist() == [4, 6]
    assert plan["vis_src"].tolist() == [3, 4, 5]          # row 1 uses the SECOND block (vtimellm_arch.py:168-176)
    assert plan["vis_dst"].tolist() == [5, 6, 7]


def test_windows_selection_and_merge_match_oracle():
    for ctx, clip, nf in ((18000, 625, 250), (7200, 1000, 100), (700, 625, 250)):
        np.testing.assert_array_equal(scoring.stage1_windows(ctx, clip, nf), scoring_ref.stage1_windows(ctx, clip, nf))
        a, ta = scoring.stage2_windows(ctx, clip, nf, 5)
        b, tb = scoring_ref.stage2_windows(ctx, clip, nf, 5)
        np.testing.assert_array_equal(a, b)
        assert ta == tb
    assert scoring.nonoverlap_segments(18000, 100).shape == (180, 100)
    answers = ["Not Present", "From 3 to 9.", "Not Present", "From 1 to 2.", "Not Present"] * 3
    for batch in (4, 10, 33):
        assert scoring.stage2_select_windows(answers, 40, batch) == scoring_ref.stage2_select_windows(answers, 40, batch)
    assert scoring.parse_span("From 34 to 12.") == (12, 34) == scoring_ref.parse_span("From 34 to 12.")
    assert scoring.parse_span("Not Present") is None
    assert scoring.parse_first_int("video 17 and 3") == 17
    cos, ent = [0.2, 0.5, 0.4], [1.0, 4.0, 2.0]
    for mode in ("add", "multiply", "neg"):
        assert scoring.merge_scores(cos, ent, mode) == scoring_ref.merge_scores(cos, ent, mode)


def test_entropy_stats_from_steps_matches_oracle():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3, 5, 97, generator=g) * 2
    ent = torch.stack([scoring_ref.step_entropy(logits[:, t]) for t in range(5)], dim=1)
    np.testing.assert_allclose(scoring.entropy_stats_from_steps(ent).numpy(),
                               scoring_ref.get_entropy_statistics(logits).numpy(), rtol=1e-6, atol=1e-6)
    one = scoring.entropy_stats_from_steps(ent[:, :1])
    assert one[:, 3].abs().max() == 0


def test_record_pack_roundtrip_and_shards():
    n = 7
    tok = torch.arange(n * 5, dtype=torch.int32).view(n, 5)
    spans = torch.tensor([[i, i + 3] for i in range(n)], dtype=torch.int32)
    hm, hx, cs = torch.rand(n), torch.rand(n) + 1, torch.rand(n) - 0.5
    rec = sweep.pack_records(tok, spans, hm, hx, cs)
    un = sweep.unpack_records(rec)
    assert rec.shape == (n, sweep.REC_WORDS) and rec.dtype == torch.int32
    assert torch.equal(un["tokens"][:, :5], tok) and (un["tokens"][:, 5:] == -1).all()
    assert torch.equal(un["spans"], spans) and torch.equal(un["h_mean"], hm) and torch.equal(un["cos"], cs)
    all_idx = np.sort(np.concatenate([sweep.shard_indices(n, r, 3) for r in range(3)]))
    assert all_idx.tolist() == list(range(n))


def test_cabi_exports_every_symbol_the_header_declares():
    header = open(os.path.join(ROOT, "include", "revisionllm_b200.h")).read()
    declared = sorted(set(re.findall(r"RVL_API[^;(]*?\b(rvl_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_cabi.PROTOTYPES) == declared, "ctypes prototypes and header are out of sync"
    m = re.search(r"#define RVL_ABI_VERSION (\d+)", header)
    assert lib.rvl_abi_version() == int(m.group(1))
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert [e for e in exported if e.startswith("rvl_")] == declared


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the loud failure on a box without a GPU")
def test_product_path_fails_loudly_without_gpu():
    lib = _cabi.load()
    h = ctypes.c_void_p()
    cfg = _cabi.rvl_config(256, 2, 2, 128, 512, 512, 768, 1024, 32, 0, 1e-5, 10000.0)
    rc = lib.rvl_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and "no CPU fallback" in _cabi.last_error()
    from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM
    model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(syn.TINY), {})
    with pytest.raises(_cabi.RvlError):
        model.cuda()
    with pytest.raises(_cabi.RvlError):
        model.generate(torch.zeros(1, 4, dtype=torch.long), images=torch.zeros(1, 2, 768))
    with pytest.raises(_cabi.RvlError):
        model.float()


def test_metrics_host_mirror_matches_oracle(golden_dir):
    import json
    from oracle import metrics_ref
    from revisionllm_b200 import metrics
    g = json.load(open(os.path.join(golden_dir, "merge_metrics.json")))
    for c in g["iou_cases"]:
        a = metrics.iou(c["answers"], tuple(c["gt"]), c["num_frames_clip"], c["num_frames_video"], c["scores"], c["plus_baseline"])
        b = metrics_ref.iou(c["answers"], tuple(c["gt"]), c["num_frames_clip"], c["num_frames_video"], c["scores"], c["plus_baseline"])
        assert a == b
    merged = [metrics_ref.merge_with_retrieval(q["gl"], q["rl"], q["rl2"]) for q in g["queries"]]
    ranked = []
    for m in merged:
        sc = m["info"]["scores"]
        order = sorted(range(len(sc)), key=lambda k: sc[k], reverse=True)
        ranked.append([m["info"]["iou"][i] for i in order])
    got = metrics.grounding_metrics_stream(ranked)
    for k, v in g["metrics"].items():
        assert abs(got[k] - v) < 1e-9
    q = g["queries"][0]
    n = len(q["gl"]["answer"])
    cov = metrics.cover_mask(n, q["rl"]["info"]["frames"])
    ref = set()
    for lo, hi in q["rl"]["info"]["frames"].values():
        ref |= set(range(max(0, int(.4 * lo)), min(int(.4 * hi), n - 1)))
    assert set(np.nonzero(cov)[0].tolist()) == ref


class _FakeStage2Model:
    """Stands in for the CUDA model in the stage-2 drivers: every prompt's new tokens are a deterministic function of its
    prompt ids (unpadded), its visual rows in order and its query features, so the drivers' batching, padding, window
    de-duplication and rank assignment can be checked on the CPU."""

    def __init__(self):
        self.calls = []

    def generate(self, ids, images=None, query_feats=None, attention_mask=None, max_new_tokens=3, **kw):
        from revisionllm_b200.model import WindowBank
        B, L = ids.shape
        self.calls.append(B)
        if isinstance(images, WindowBank):
            per_window = images.windows.float().sum(dim=(1, 2))                       # [U]
            qsum = (query_feats[0].float() * query_feats[1][..., None].float()).sum(dim=(1, 2))   # [Q]
            per_window = per_window + 7.0 * qsum[images.text_index.long()]
            rows = list(images.rows) if isinstance(images.rows, (list, tuple)) else list(images.rows.unbind(0))
            vis = [per_window[r] for r in rows]                                         # B x [V_b]
        else:
            v4 = images.float().sum(dim=(2, 3))
            if query_feats is not None:
                v4 = v4 + 7.0 * (query_feats[0].float() * query_feats[1][..., None].float()).sum(dim=(1, 2))[:, None]
            vis = list(v4.unbind(0))
        am = torch.ones_like(ids, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
        code = torch.stack([(v * torch.arange(1, v.shape[0] + 1, dtype=torch.float32)).sum() for v in vis]) + \
            (ids.clamp(min=0) * am).sum(dim=1).float()
        new = torch.stack([(code * (t + 1)).round().long() % 97 for t in range(max_new_tokens)], dim=1)
        ent = torch.rand(B, max_new_tokens, generator=torch.Generator().manual_seed(0)) + 0.5
        return {"sequences": torch.cat([ids, new], dim=1), "entropies": ent}


def test_stage2_drivers_batch_pad_dedup_and_shard_like_the_sequential_schedule():
    g = torch.Generator().manual_seed(5)
    queries = []
    for k, (nw, L, lq) in enumerate([(7, 14, 5), (5, 11, 3), (9, 14, 5), (4, 9, 2), (3, 12, 4)]):   # 3 windows: a 3-row prompt among 4-row ones
        ids = torch.randint(3, 500, (L,), generator=g)
        ids[4] = -200
        queries.append(dict(windows=torch.randint(-3, 4, (nw, 6, 8), generator=g).float(),
                            query_feats=(torch.randint(-2, 3, (1, lq, 8), generator=g).float(), torch.ones(1, lq)),
                            input_ids=ids, grounding_windows=list(range(100 * k, 100 * k + nw)), perm_seed=k))
    kw = dict(batch=4, zooms=(2, 1), max_new_tokens=3, answer_number=lambda t: int(t[0]) % 4, eos_token_id=None)
    key = lambda res: [(r["zoom"], r["start"], r["perm"], r["tokens"], r["window"]) for r in res]
    # the reference's schedule: one generate() per chunk, stacked zoom repeats
    seq_model = _FakeStage2Model()
    seq = [sweep.stage2_pass(seq_model, q["windows"], q["query_feats"], q["input_ids"], q["grounding_windows"], perm_seed=k,
                             max_calls_per_batch=1, dedup=False, **kw) for k, q in enumerate(queries)]
    assert all(b == 1 for b in seq_model.calls)
    # batched across queries, windows de-duplicated, prompts and query features of different lengths padded + masked
    multi_model = _FakeStage2Model()
    multi = sweep.stage2_pass_queries(multi_model, queries, **kw)
    assert [key(r) for r in multi] == [key(r) for r in seq]
    assert sum(multi_model.calls) == sum(seq_model.calls) and len(multi_model.calls) < len(seq_model.calls)
    # stacked copies instead of the bank: same answers
    stacked = sweep.stage2_pass_queries(_FakeStage2Model(), queries, dedup=False, **kw)
    assert [key(r) for r in stacked] == [key(r) for r in seq]
    # a batch bound splits the work, query i belongs to rank i mod world
    small = sweep.stage2_pass_queries(_FakeStage2Model(), queries, max_calls_per_batch=3, **kw)
    assert [key(r) for r in small] == [key(r) for r in seq]
    for rank in range(2):
        part = sweep.stage2_pass_queries(_FakeStage2Model(), queries, rank=rank, world=2, **kw)
        for qi in range(len(queries)):
            assert (part[qi] is None) == (qi % 2 != rank)
            if part[qi] is not None:
                assert key(part[qi]) == key(seq[qi])
    # every window picked maps back into its own query's grounding windows
    for k, res in enumerate(multi):
        assert all(r["window"] in queries[k]["grounding_windows"] for r in res)


class _FakeSweepModel:
    """generate() for the ragged-sweep driver on the CPU: new tokens are a function of each window's content and prompt."""
    engine = None
    device = torch.device("cpu")

    def __init__(self):
        self.batches = []

    def generate(self, ids, images=None, attention_mask=None, max_new_tokens=4, **kw):
        am = attention_mask.bool()
        self.batches.append(sum(int(am[i].sum()) - 1 + int(images[i].shape[0]) for i in range(ids.shape[0])))
        code = torch.stack([im.float().sum() for im in images]) + (ids.clamp(min=0) * am).sum(dim=1).float()
        new = torch.stack([(code * (t + 1)).round().long() % 89 for t in range(max_new_tokens)], dim=1)
        ent = torch.full((ids.shape[0], max_new_tokens), 0.25)
        return {"sequences": torch.cat([ids, new], dim=1), "entropies": ent}


def test_ragged_sweep_packs_by_token_budget_and_keeps_window_order():
    g = torch.Generator().manual_seed(9)
    n = 23
    frames = [int(x) for x in torch.randint(1, 60, (n,), generator=g)]
    windows = [torch.randint(-3, 4, (f, 8), generator=g).float() for f in frames]
    L = 12
    ids = torch.randint(3, 300, (n, L), generator=g)
    ids[:, 3] = -200
    am = torch.ones(n, L, dtype=torch.bool)
    for i in range(n):
        am[i, L - (i % 4):] = False if i % 4 else am[i, L:]          # right padding of 0..3 positions
    budget = 150
    model = _FakeSweepModel()
    rec = sweep.ragged_sweep(model, windows, ids, am, None, max_new_tokens=4, rank=0, world=1, max_tokens_per_batch=budget, eos_token_id=None)
    got = sweep.unpack_records(rec)["tokens"][:, :4]
    # one window at a time gives the same tokens, in window order
    for i in range(n):
        one = _FakeSweepModel().generate(ids[i:i + 1], images=[windows[i]], attention_mask=am[i:i + 1], max_new_tokens=4)
        assert got[i].tolist() == one["sequences"][0, L:].tolist(), i
    # every packed batch respects the token budget unless it holds a single over-long window
    assert len(model.batches) > 1 and all(b <= budget for b in model.batches)
    assert sum(model.batches) == sum(int(am[i].sum()) - 1 + frames[i] for i in range(n))


def test_shared_prefix_plan_remap_is_a_compact_relabelling():
    """model._share_prefix_rows: the packed stream becomes [prefix | seq 0 from P | seq 1 from P | ...].  Checked on the plan
    alone: every destination row is used exactly once, the prefix rows carry sequence 0's first P items, every other
    item keeps its order inside its sequence, and nothing but the duplicated prefix items disappears."""
    from types import SimpleNamespace
    from revisionllm_b200.model import RevisionLlamaForCausalLM as M
    B, F, ps = 5, 40, 32
    g = np.random.default_rng(3)
    pre = g.integers(3, 900, size=37)                                   # common system prompt: 37 ids -> P = 32
    ids = np.stack([np.concatenate([pre, [-200], g.integers(3, 900, size=9 + 0)]) for _ in range(B)]).astype(np.int64)
    n_vis = [F - 3 * i for i in range(B)]                               # ragged visual blocks
    plan = plan_splice(ids, n_vis)
    plan["shared_prefix"] = min(M._common_text_prefix(ids, None), int(plan["lengths"].min()))
    plan["ctx_len"] = 0
    assert plan["shared_prefix"] == 37
    before = {k: plan[k].copy() for k in ("cu_seqlens", "text_ids", "text_dst", "vis_src", "vis_dst", "lengths")}
    M._share_prefix_rows(SimpleNamespace(engine=SimpleNamespace(cfg=SimpleNamespace(kv_page_size=ps))), plan)
    P = plan["ctx_len"]
    assert P == 32
    cu = plan["cu_seqlens"]
    lengths = before["lengths"].astype(np.int64)
    assert cu.tolist() == np.concatenate([[0, P], P + np.cumsum(lengths - P)]).tolist()      # B + 1 sequences, the prefix first
    dst = np.concatenate([plan["text_dst"], plan["vis_dst"]])
    assert sorted(dst.tolist()) == list(range(int(cu[-1])))                                    # every row written exactly once
    # prefix rows = the first P text ids of sequence 0, in order
    order = np.argsort(plan["text_dst"])
    assert plan["text_ids"][order][:P].tolist() == pre[:P].tolist()
    # each sequence keeps its items from position P on, in order: compare with the unshared plan
    ocu = before["cu_seqlens"].astype(np.int64)
    for b in range(B):
        lo, hi = ocu[b] + P, ocu[b + 1]
        old_t = [(d - lo, t) for d, t in zip(before["text_dst"], before["text_ids"]) if lo <= d < hi]
        new_t = [(d - cu[b + 1], t) for d, t in zip(plan["text_dst"], plan["text_ids"]) if cu[b + 1] <= d < cu[b + 2]]
        assert sorted(old_t) == sorted(new_t), b
        old_v = [(d - lo, s) for d, s in zip(before["vis_dst"], before["vis_src"]) if lo <= d < hi]
        new_v = [(d - cu[b + 1], s) for d, s in zip(plan["vis_dst"], plan["vis_src"]) if cu[b + 1] <= d < cu[b + 2]]
        assert sorted(old_v) == sorted(new_v), b
    # a batch of one, or a common prefix shorter than a page, is left alone
    small = plan_splice(ids[:, 10:], n_vis)
    small["shared_prefix"], small["ctx_len"] = 27, 0
    keep = small["cu_seqlens"].copy()
    M._share_prefix_rows(SimpleNamespace(engine=SimpleNamespace(cfg=SimpleNamespace(kv_page_size=ps))), small)
    assert small["ctx_len"] == 0 and np.array_equal(small["cu_seqlens"], keep)


def test_shared_context_plan_with_groups_of_prompts():
    """model._share_prefix_rows with `plan["group"]` (several prompts per segment): one context sequence per group in front
    of the packed stream, carrying the group leader's first P items; every prompt keeps its items from position P on; every row
    of the new stream is written exactly once.  And the page table gives each group its own copy of the shared pages."""
    from types import SimpleNamespace
    from revisionllm_b200.model import RevisionLlamaForCausalLM as M
    S, Q, F, ps, n_pre = 3, 4, 50, 32, 20
    g = np.random.default_rng(7)
    pre = g.integers(3, 900, size=n_pre)
    ids = np.stack([np.concatenate([pre, [-200], g.integers(3, 900, size=9)]) for _ in range(S * Q)]).astype(np.int64)
    group = np.repeat(np.array([7, 2, 5]), Q)                            # arbitrary labels, segment-major rows
    plan = plan_splice(ids, [F] * (S * Q))
    plan["shared_prefix"], plan["ctx_len"], plan["group"] = n_pre + F, 0, group
    before = {k: plan[k].copy() for k in ("cu_seqlens", "text_ids", "text_dst", "vis_src", "vis_dst", "lengths")}
    M._share_prefix_rows(SimpleNamespace(engine=SimpleNamespace(cfg=SimpleNamespace(kv_page_size=ps))), plan)
    P, G = plan["ctx_len"], plan["ctx_groups"]
    assert (P, G) == (64, 3)
    lengths = before["lengths"].astype(np.int64)
    cu = plan["cu_seqlens"]
    assert cu.tolist() == np.concatenate([np.arange(G) * P, G * P + np.concatenate([[0], np.cumsum(lengths - P)])]).tolist()
    dst = np.concatenate([plan["text_dst"], plan["vis_dst"]])
    assert sorted(dst.tolist()) == list(range(int(cu[-1])))
    assert plan["group_of"].tolist() == np.repeat([2, 0, 1], Q).tolist() and sorted(plan["group_leader"].tolist()) == [0, 4, 8]
    ocu = before["cu_seqlens"].astype(np.int64)
    for gi in range(G):                                                 # context gi = the first P items of its leader
        b = int(plan["group_leader"][gi])
        old = sorted([(d - ocu[b], "t", t) for d, t in zip(before["text_dst"], before["text_ids"]) if ocu[b] <= d < ocu[b] + P] +
                     [(d - ocu[b], "v", v) for d, v in zip(before["vis_dst"], before["vis_src"]) if ocu[b] <= d < ocu[b] + P])
        new = sorted([(d - cu[gi], "t", t) for d, t in zip(plan["text_dst"], plan["text_ids"]) if cu[gi] <= d < cu[gi + 1]] +
                     [(d - cu[gi], "v", v) for d, v in zip(plan["vis_dst"], plan["vis_src"]) if cu[gi] <= d < cu[gi + 1]])
        assert old == new, gi
    for b in range(S * Q):                                              # own rows: the items from position P on, in order
        lo, hi = ocu[b] + P, ocu[b + 1]
        old = sorted([(d - lo, t) for d, t in zip(before["text_dst"], before["text_ids"]) if lo <= d < hi] +
                     [(d - lo, -1 - v) for d, v in zip(before["vis_dst"], before["vis_src"]) if lo <= d < hi])
        new = sorted([(d - cu[G + b], t) for d, t in zip(plan["text_dst"], plan["text_ids"]) if cu[G + b] <= d < cu[G + b + 1]] +
                     [(d - cu[G + b], -1 - v) for d, v in zip(plan["vis_dst"], plan["vis_src"]) if cu[G + b] <= d < cu[G + b + 1]])
        assert old == new, b
    # one prompt per group: nothing to share
    solo = plan_splice(ids[:3], [F] * 3)
    solo["shared_prefix"], solo["ctx_len"], solo["group"] = n_pre + F, 0, np.array([0, 1, 2])
    M._share_prefix_rows(SimpleNamespace(engine=SimpleNamespace(cfg=SimpleNamespace(kv_page_size=ps))), solo)
    assert solo["ctx_len"] == 0
    # page table: two shared pages per group, disjoint between groups, own pages disjoint everywhere
    asked = []
    stub = SimpleNamespace(engine=SimpleNamespace(cfg=SimpleNamespace(kv_page_size=ps), ensure_kv=asked.append),
                           device=torch.device("cpu"), share_prefix_pages=True)
    kv = M._alloc_kv(stub, lengths, extra=16, shared_prefix=n_pre + F, group=group)
    t = kv.page_table.numpy()
    need = int(np.ceil((int(lengths[0]) + 16) / ps))
    for gi in range(S):
        rows = t[gi * Q:(gi + 1) * Q]
        assert (rows[:, :2] == rows[0, :2]).all()
    shared = {tuple(r[:2]) for r in t}
    assert len(shared) == S and len({p for pair in shared for p in pair}) == 2 * S
    own = t[:, 2:need].reshape(-1).tolist()
    assert len(own) == len(set(own)) and not set(own) & {p for pair in shared for p in pair}
    assert asked == [2 * S + len(own)]


def test_kv_page_table_shares_whole_prefix_pages_only():
    """model._alloc_kv: pages filled entirely by the common prompt prefix are mapped once for all sequences; every other page
    belongs to exactly one sequence; every sequence can hold its prompt plus the reserved new tokens."""
    from types import SimpleNamespace
    from revisionllm_b200.model import RevisionLlamaForCausalLM as M
    ps = 32
    asked = []
    stub = SimpleNamespace(engine=SimpleNamespace(cfg=SimpleNamespace(kv_page_size=ps), ensure_kv=asked.append),
                           device=torch.device("cpu"), share_prefix_pages=True)
    lengths = np.array([184, 184, 150, 97, 70], dtype=np.int64)
    kv = M._alloc_kv(stub, lengths, extra=16, shared_prefix=70)          # 70 common positions -> 2 whole pages
    table = kv.page_table.numpy()
    need = [int(np.ceil((l + 16) / ps)) for l in lengths]
    assert table.shape == (5, max(need))
    assert (table[:, :2] == table[0, :2]).all() and table[0, :2].tolist() == [0, 1]
    own = [table[i, 2:need[i]].tolist() for i in range(5)]
    flat = [p for row in own for p in row]
    assert len(flat) == len(set(flat)) and not set(flat) & {0, 1}        # disjoint, and never a shared page
    assert asked == [2 + len(flat)] and max(flat) == asked[0] - 1         # the pool is asked for exactly what is mapped
    # a prefix longer than the shortest prompt shares only that prompt's whole pages; one sequence shares nothing
    kv = M._alloc_kv(stub, np.array([184, 40], dtype=np.int64), extra=16, shared_prefix=100)
    t = kv.page_table.numpy()
    assert (t[:, 0] == 0).all() and t[0, 1] != t[1, 1]
    kv = M._alloc_kv(stub, np.array([184], dtype=np.int64), extra=16, shared_prefix=100)
    assert kv.page_table.numpy()[0].tolist() == list(range(int(np.ceil(200 / ps))))
    stub.share_prefix_pages = False
    t = M._alloc_kv(stub, lengths, extra=16, shared_prefix=70).page_table.numpy()
    assert len({int(p) for i in range(5) for p in t[i, :need[i]]}) == sum(need)


def test_checkpoint_readers_and_key_cleanup(tmp_path):
    """checkpoint.load_hf_checkpoint over every directory layout builder.py's `from_pretrained` accepts (sharded / single,
    safetensors / .bin) and the key clean-up of builder.py:12-15."""
    import json
    from safetensors.torch import save_file
    from revisionllm_b200 import checkpoint
    from revisionllm_b200._cabi import RvlError
    g = torch.Generator().manual_seed(1)
    sd = {f"model.layers.{i}.mlp.up_proj.weight": torch.randn(4, 3, generator=g).to(torch.bfloat16) for i in range(4)}
    sd["lm_head.weight"] = torch.randn(5, 3, generator=g).to(torch.bfloat16)
    keys = sorted(sd)
    same = lambda a: set(a) == set(sd) and all(torch.equal(a[k], sd[k]) for k in sd)
    # sharded safetensors with an index
    d = tmp_path / "st_sharded"; os.makedirs(d)
    parts = {"model-00001-of-00002.safetensors": keys[:2], "model-00002-of-00002.safetensors": keys[2:]}
    for f, ks in parts.items():
        save_file({k: sd[k] for k in ks}, str(d / f))
    json.dump({"weight_map": {k: f for f, ks in parts.items() for k in ks}}, open(d / "model.safetensors.index.json", "w"))
    assert same(checkpoint.load_hf_checkpoint(str(d)))
    # sharded .bin with an index
    d = tmp_path / "bin_sharded"; os.makedirs(d)
    parts = {"pytorch_model-00001-of-00002.bin": keys[:3], "pytorch_model-00002-of-00002.bin": keys[3:]}
    for f, ks in parts.items():
        torch.save({k: sd[k] for k in ks}, d / f)
    json.dump({"weight_map": {k: f for f, ks in parts.items() for k in ks}}, open(d / "pytorch_model.bin.index.json", "w"))
    assert same(checkpoint.load_hf_checkpoint(str(d)))
    # single files
    d = tmp_path / "st_single"; os.makedirs(d)
    save_file(sd, str(d / "model.safetensors"))
    assert same(checkpoint.load_hf_checkpoint(str(d)))
    d = tmp_path / "bin_single"; os.makedirs(d)
    torch.save(sd, d / "pytorch_model.bin")
    assert same(checkpoint.load_hf_checkpoint(str(d)))
    d = tmp_path / "empty"; os.makedirs(d)
    with pytest.raises(RvlError):
        checkpoint.load_hf_checkpoint(str(d))
    # builder.py:12-15: strip "base_model." and, if any key then starts with "model.model.", one more "model."
    raw = {"base_model.model.model.norm.weight": 1, "base_model.model.model.mm_projector.bias": 2, "base_model.model.lm_head.weight": 3}
    assert checkpoint.strip_lora_prefixes(raw) == {"model.norm.weight": 1, "model.mm_projector.bias": 2, "lm_head.weight": 3}
    assert checkpoint.strip_lora_prefixes({"model.norm.weight": 1}) == {"model.norm.weight": 1}


def test_feature_store_reads_blobs_written_by_the_reference_writer(golden_dir, tmp_path):
    """features.FeatureStore against LMDB values produced by the reference's own `dumps_npz` (tests/golden/make_golden_blobs.py),
    its own writer, a directory of .npy files, and the failure modes."""
    from revisionllm_b200 import features as ft
    from revisionllm_b200._cabi import RvlError
    g = np.load(os.path.join(golden_dir, "feature_blobs.npz"))
    store = ft.FeatureStore({"movie": g["video_blob"].tobytes(), "q": g["query_blob"].tobytes()})
    assert store.kind == "dict"
    v = store.video("movie")
    assert v.dtype == np.float32 and np.array_equal(v, g["features"])
    tok, cls = store.query("q")
    assert np.array_equal(tok, g["token_features"]) and np.array_equal(cls, g["cls_features"])
    with pytest.raises(KeyError):
        store.video("absent")
    # our writer produces blobs the same reader (np.load on the bytes, eval_nlq_negative.py:193-197) accepts
    again = ft.loads_npz(ft.dumps_npz({"features": g["features"]}))
    assert np.array_equal(again["features"], g["features"])
    # bare .npy per video (VidChapters, eval_nlq_negative.py:198-200)
    np.save(tmp_path / "vid7.npy", g["features"])
    npy = ft.FeatureStore(str(tmp_path))
    assert npy.kind == "npy" and np.array_equal(npy.video("vid7"), g["features"])
    # an LMDB environment needs the `lmdb` module, which this image does not have: loud failure, no fallback
    try:
        import lmdb  # noqa: F401
    except ImportError:
        os.makedirs(tmp_path / "env")
        with pytest.raises(RvlError):
            ft.FeatureStore(str(tmp_path / "env"))


def test_window_builders_match_tables_produced_by_the_reference_source(golden_dir):
    """tests/golden/windows.json holds what the reference's own window-building source lines (eval_nlq_negative.py:224-240,
    eval_nlq_retrieval_e2e2.py:262-294, exec'd by make_golden_windows.py) produce; the host mirrors and the oracle must agree."""
    import json
    g = json.load(open(os.path.join(golden_dir, "windows.json")))
    for c in g["stage1"]:
        clip_length = c["debug_window"] * c["feature_fps"]
        for impl in (scoring.stage1_windows, scoring_ref.stage1_windows):
            idx = impl(c["ctx_l"], clip_length, c["num_frames"])
            rows = idx.tolist()
            if c["baseline"]:
                rows = [rows[1]]                                          # --baseline keeps window 1 only (:228)
            if c["plus_baseline"]:                                        # the whole-video row of --plus_baseline (:237-240)
                rows = rows + [np.linspace(0, c["ctx_l"] - 1, c["num_frames"], dtype=np.int32).tolist()]
            assert rows == c["windows"], (impl.__module__, c["ctx_l"], clip_length)
    for c in g["stage2"]:
        clip_length = c["debug_window"] * c["feature_fps"]
        for mod in (scoring, scoring_ref):
            idx, times = mod.stage2_windows(c["ctx_l"], clip_length, c["num_frames"], c["stride"])
            assert [list(t) for t in times] == c["times"] and idx.shape[0] == c["n_windows"]
            if c["stage1_answers"] is None:
                sel = list(range(idx.shape[0]))                           # no stage-1 log: every window (:293-294)
            else:
                sel = mod.stage2_select_windows(c["stage1_answers"], idx.shape[0], c["batch"], c["stride"])
            want = c["grounding_windows"]
            # exact, order included: sorted after padding, CPython's set order of the mapped ids otherwise (:284-290)
            assert sel == want, (mod.__name__, c["ctx_l"], c["batch"])
            # negative ids (a positive stage-1 window 0 maps to -3..-1) index from the end, like the reference's list indexing
            first = [int(idx[i][0]) for i in want]
            assert first == c["selected_first_frames"]


def test_stage2_call_plan_matches_the_reference_loop(golden_dir):
    """sweep._stage2_plan against tests/golden/stage2_plan.json (the reference's own zoom loop, exec'd by
    make_golden_stage2_plan.py): chunk starts, permutations (same generator seed) and the windows of every call, including
    the cases with fewer windows than one chunk, where the reference's negative slice start applies."""
    import json
    g = json.load(open(os.path.join(golden_dir, "stage2_plan.json")))
    for c in g["cases"]:
        gen = torch.Generator().manual_seed(c["seed"])
        plan = sweep._stage2_plan(c["n_windows"], c["batch"], (4, 2, 1), gen)
        assert [p["zoom"] for p in plan] == c["zooms"]
        assert [p["start"] for p in plan] == c["starts"], (c["n_windows"], c["batch"])
        assert [p["idx"].tolist() for p in plan] == c["perms"]
        assert [p["rows"].tolist() for p in plan] == c["call_windows"]


def test_stage2_frames_match_the_reference_iou(golden_dir):
    """metrics.stage2_frames and the oracle's stage2_iou against tests/golden/stage2_iou.json (the stage-2 script's own `iou`,
    lifted by make_golden_stage2_iou.py); the window `sweep._stage2_finish` reports is the one those frames are built around."""
    import json
    from oracle import metrics_ref
    from revisionllm_b200 import metrics
    g = json.load(open(os.path.join(golden_dir, "stage2_iou.json")))
    assert any(c["hit"] == [1] for c in g["cases"]) and any(c["hit"] == [0] for c in g["cases"])
    for c in g["cases"]:
        args = (c["outputs"], c["gt"], c["num_frames_video"], c["starts"], c["indexes"], c["hierarchy_zooms"], c["grounding_windows"])
        want = {int(k): tuple(v) for k, v in c["clip_frames"].items()}
        for fn in (metrics.stage2_frames, metrics_ref.stage2_iou):
            frames, hit = fn(*args)
            assert {k: tuple(v) for k, v in frames.items()} == want and hit == c["hit"], fn.__module__
        # the driver's own mapping (first integer // zoom -> permuted chunk -> grounding window)
        calls = [dict(zoom=z, start=s0, n=len(ix), idx=torch.tensor(ix)) for z, s0, ix in zip(c["hierarchy_zooms"], c["starts"], c["indexes"])]
        results = [dict(tokens=torch.tensor([0]), stats=torch.tensor([1.0, 1.0, 1.0, 0.0])) for _ in calls]
        texts = iter(c["outputs"])

        def number(_tok):
            m = re.search(r"(\d+)", next(texts))
            return int(m.group(1)) if m else None
        out = sweep._stage2_finish(calls, results, c["grounding_windows"], number)
        for i, r in enumerate(out):
            if i in want:
                w = r["window"]
                assert (max(0, w - 1), min(c["num_frames_video"], w + 1)) == want[i]
            else:
                assert r["window"] is None


def test_inline_merge_matches_the_reference_source(golden_dir):
    """scoring.merge_scores (and the oracle's) against tests/golden/merge_inline.json: the reference's own normalise + merge
    lines (eval_nlq_negative.py:321-335), exec'd by make_golden_merge.py.  Bit-exact: the same Python float arithmetic."""
    import json
    g = json.load(open(os.path.join(golden_dir, "merge_inline.json")))
    for c in g["cases"]:
        for mod in (scoring, scoring_ref):
            if c["score"] == "cosine_sim":                         # 'entropy' not in args.score: the (normalised) cosine score alone
                cos = c["score_cos"]
                got = [v / max(cos) for v in cos] if (c["normalize"] and cos) else list(cos)
            else:
                got = mod.merge_scores(c["score_cos"], c["scores_entropy"], mode=c["score_merge"], normalize=c["normalize"])
            assert got == c["scores"], (mod.__name__, c["score"], c["score_merge"], c["normalize"], len(c["score_cos"]))


def test_every_entry_point_rejects_a_null_handle_without_a_gpu():
    """Error behaviour of the C ABI (include/revisionllm_b200.h): a call with a NULL handle / NULL buffers returns a negative
    rvl_status and leaves a message in rvl_last_error - no crash, no CUDA call, so this runs on the CPU-only box."""
    lib = _cabi.load()
    skipped = {"rvl_abi_version", "rvl_last_error", "rvl_destroy", "rvl_create", "rvl_workspace_bytes", "rvl_kv_bytes",
               "rvl_debug_gemm_timestamps", "rvl_reload_env", "rvl_debug_sm_clock", "rvl_debug_attn_timestamps", "rvl_clip_encoder_workspace_bytes"}
    checked = 0
    for name, (restype, argtypes) in _cabi.PROTOTYPES.items():
        if name in skipped:
            continue
        args = []
        for t in argtypes:
            if t is ctypes.c_void_p or hasattr(t, "contents") or (hasattr(t, "_type_") and not isinstance(t._type_, str)):
                args.append(None)
            elif t in (ctypes.c_float, ctypes.c_double):
                args.append(t(0.0))
            else:
                args.append(t(0))
        rc = getattr(lib, name)(*args)
        assert isinstance(rc, int) and rc < 0, (name, rc)
        msg = lib.rvl_last_error(None)
        assert msg and name.encode() in msg, (name, msg)
        checked += 1
    assert checked >= 20
    assert lib.rvl_create(None, None) < 0
    lib.rvl_destroy(None)                                  # a no-op, not a crash


def test_rvl_create_validates_the_configuration_before_touching_cuda():
    """include/revisionllm_b200.h `rvl_create`: configurations the kernels do not cover are refused with a message naming the
    reason (checked before any CUDA call, so the CPU-only box sees them); a valid configuration then fails loudly for want of
    a GPU - there is no CPU fallback."""
    lib = _cabi.load()
    good = dict(hidden=4096, n_layers=32, n_heads=32, head_dim=128, intermediate=11008, vocab=32000, adapter_dim=768, max_pos=4096,
                kv_page_size=32, device=0, rms_eps=1e-5, rope_theta=10000.0)

    def create(**over):
        cfg = _cabi.rvl_config(**{**good, **over})
        h = ctypes.c_void_p()
        rc = lib.rvl_create(ctypes.byref(cfg), ctypes.byref(h))
        return rc, (lib.rvl_last_error(None) or b"").decode(), h
    for over, needle in ((dict(head_dim=64, hidden=2048), "head_dim must be 128"), (dict(hidden=4224), "hidden != n_heads*head_dim"),
                         (dict(intermediate=11000), "unsupported dimensions"), (dict(vocab=32001), "unsupported dimensions"),
                         (dict(adapter_dim=770), "unsupported dimensions"), (dict(kv_page_size=0), "bad kv_page_size"),
                         (dict(kv_page_size=512), "bad kv_page_size")):
        rc, msg, h = create(**over)
        assert rc < 0 and needle in msg and not h.value, (over, rc, msg)
    if not torch.cuda.is_available():
        rc, msg, h = create()
        assert rc < 0 and "no CUDA device" in msg and "no CPU fallback" in msg and not h.value


def test_plan_splice_equals_oracle_splice_on_random_ragged_batches():
    """engine.plan_splice (the product's host index plan, vectorised fast path and general loop) replayed with plain gathers
    against oracle/splice_ref.splice over 60 random batches: masks, rows without a placeholder, frame counts from 0 to 9,
    truncation inside the text in front of, inside, and behind the visual block."""
    rng = np.random.default_rng(11)
    H, V = 16, 64
    w = {"model.embed_tokens.weight": torch.from_numpy(rng.standard_normal((V, H)).astype(np.float32))}
    for case in range(60):
        B = int(rng.integers(1, 6))
        Ltxt = int(rng.integers(4, 12))
        ids = rng.integers(3, V, size=(B, Ltxt)).astype(np.int64)
        has = rng.random(B) < (0.75 if case % 3 else 1.0)                  # every third case: all rows have a placeholder (fast path)
        for b in range(B):
            if has[b]:
                ids[b, int(rng.integers(0, Ltxt))] = -200
        mask = None
        if case % 2:
            mask = np.ones((B, Ltxt), dtype=bool)
            for b in range(B):
                cut = int(rng.integers(0, 3))
                if cut:
                    keep = ids[b] == -200
                    mask[b, Ltxt - cut:] = keep[Ltxt - cut:]                 # never mask the placeholder itself away
        frames = [int(rng.integers(0 if case % 5 == 0 else 1, 10)) for _ in range(B)]
        max_len = None if case % 4 else int(rng.integers(2, Ltxt + 6))
        proj = [torch.from_numpy(rng.standard_normal((f, H)).astype(np.float32)) for f in frames]
        want = splice_ref.splice(w, torch.from_numpy(ids), proj, attention_mask=None if mask is None else torch.from_numpy(mask), max_length=max_len)
        plan = plan_splice(ids, frames, mask, max_length=max_len)
        assert plan["lengths"].tolist() == [e.shape[0] for e in want], case
        allproj = torch.cat(proj) if sum(frames) else torch.zeros(0, H)
        packed = torch.full((int(plan["cu_seqlens"][-1]), H), float("nan"))
        packed[torch.from_numpy(plan["text_dst"]).long()] = w["model.embed_tokens.weight"][torch.from_numpy(plan["text_ids"]).long()]
        packed[torch.from_numpy(plan["vis_dst"]).long()] = allproj[torch.from_numpy(plan["vis_src"]).long()]
        cu = plan["cu_seqlens"]
        for b in range(B):
            assert torch.equal(packed[cu[b]:cu[b + 1]], want[b]), (case, b)


def test_window_and_selection_mirrors_equal_the_oracle_on_random_inputs():
    """Host mirrors (revisionllm_b200.scoring) against the oracle restatements over 200 random configurations each - lengths
    shorter than one window, windows that do not divide the video, every stride, batches below / at / above the number of
    positive windows (the slice-step-0 case the reference would crash on is skipped: step = int(rest / missing) == 0)."""
    rng = np.random.default_rng(17)
    for _ in range(200):
        ctx = int(rng.integers(2, 20000))
        clip = int(rng.integers(2, 1500))
        nf = int(rng.integers(1, 260))
        stride = int(rng.integers(1, 7))
        if clip // 2 == 0 or clip // stride == 0:
            continue
        np.testing.assert_array_equal(scoring.stage1_windows(ctx, clip, nf), scoring_ref.stage1_windows(ctx, clip, nf))
        a, ta = scoring.stage2_windows(ctx, clip, nf, stride)
        b, tb = scoring_ref.stage2_windows(ctx, clip, nf, stride)
        np.testing.assert_array_equal(a, b)
        assert [tuple(t) for t in ta] == [tuple(t) for t in tb]
    for _ in range(200):
        n1 = int(rng.integers(1, 80))
        stride = int(rng.integers(2, 7))
        n2 = int(rng.integers(1, 200))
        batch = int(rng.integers(1, 120))
        p = float(rng.uniform(0, 1))
        answers = ["From 1 to 2" if rng.random() < p else "Not Present" for _ in range(n1)]
        got = scoring.stage2_select_windows(answers, n2, batch, stride)
        want = scoring_ref.stage2_select_windows(answers, n2, batch, stride)
        assert got == want, (n1, stride, n2, batch)
    for _ in range(100):
        n = int(rng.integers(1, 40))
        cos = [float(v) for v in rng.uniform(0.01, 1, size=n)]
        ent = [float(v) for v in rng.uniform(0.05, 4, size=n)]
        for mode in ("add", "multiply", "neg"):
            for normalize in (True, False):
                assert scoring.merge_scores(cos, ent, mode, normalize) == scoring_ref.merge_scores(cos, ent, mode, normalize)


def test_placeholder_tokenisation_matches_reference_on_more_prompts(golden_dir):
    """mm_utils.tokenizer_image_token against tests/golden/prompt_more.json (the reference's own function, run by
    make_golden_prompts.py): several <video> placeholders, <memory> with and without trailing text, tokenizers with and
    without BOS."""
    import json

    class NoBos(syn.StubTokenizer):
        def __call__(self, text):
            r = super().__call__(text)
            r.input_ids = r.input_ids[1:]
            return r
    g = json.load(open(os.path.join(golden_dir, "prompt_more.json")))
    assert len(g["cases"]) >= 12
    for c in g["cases"]:
        tok = syn.StubTokenizer(32000) if c["bos"] else NoBos(32000)
        assert mm_utils.tokenizer_image_token(c["prompt"], tok, -200) == c["ids"], (c["prompt"], c["bos"])
        assert mm_utils.tokenizer_image_token(c["prompt"], tok, -200, return_tensors="pt").tolist() == c["ids"]


def test_inference_surface_matches_the_reference_function(golden_dir):
    """revisionllm_b200.inference.inference against tests/golden/inference_surface.json (the reference's own `inference()`,
    lifted by make_golden_inference.py, run with a recording fake model): same ids and batch handed to generate(), same
    keyword surface (sampling is opt-in here - greedy is the north star - and `do_sample=True` restores the reference's
    arguments), same decoded / stripped / stop-string-trimmed answers, same list-vs-string rule, `<memory>` appended to the
    query when a visual memory is given."""
    import json
    from revisionllm_b200.inference import inference

    class FakeModel:
        def __init__(self, new_tokens):
            self.new_tokens = torch.tensor(new_tokens)

        def generate(self, input_ids, **kw):
            self.seen = (input_ids.cpu().tolist(), list(kw["images"].shape), kw)
            return {"sequences": torch.cat([input_ids.cpu(), self.new_tokens], dim=1), "scores": ()}
    g = json.load(open(os.path.join(golden_dir, "inference_surface.json")))
    tok = syn.StubTokenizer(32000)
    tok.inv[tok._word("</s>")] = "</s>"
    assert any(isinstance(c["outputs"], str) for c in g["cases"]) and any("</s>" not in str(c["outputs"]) for c in g["cases"])
    for c in g["cases"]:
        model = FakeModel(c["new_tokens"])
        B = c["batch"]
        vm = torch.zeros(B, 2, 8) if c["memory"] else None
        pm = torch.zeros(B, 3, dtype=torch.long) if c["memory"] else None
        out, mo = inference(model, torch.zeros(B, 4, 8), None, c["query"], tok, visual_memory=vm, prefix_memory=pm,
                            return_list=c["return_list"], do_sample=True)
        assert out == c["outputs"], c["query"]
        ids, shape, kw = model.seen
        assert ids == c["generate"]["input_ids"] and shape == c["generate"]["images_shape"]
        assert mo["sequences"].tolist() == c["sequences"]
        want = c["generate"]["kwargs"]
        for k in ("do_sample", "temperature", "num_beams", "max_new_tokens", "use_cache", "output_scores", "return_dict_in_generate"):
            assert kw[k] == want[k], k
        assert (kw.get("visual_memory") is not None) == c["generate"]["has_visual_memory"]
        # the default call decodes greedily with everything else unchanged
        inference(model, torch.zeros(B, 4, 8), None, c["query"], tok, visual_memory=vm, prefix_memory=pm, return_list=c["return_list"])
        assert model.seen[2]["do_sample"] is False and model.seen[0] == ids


def test_shard_balanced_invariants_on_random_lengths():
    """sweep.shard_balanced (greedy length-balanced bin packing of ragged windows onto ranks): every index exactly once,
    deterministic, no rank above the greedy bound (mean + longest item), world = 1 keeps the order of a full sort."""
    rng = np.random.default_rng(23)
    for _ in range(100):
        n = int(rng.integers(0, 400))
        world = int(rng.integers(1, 9))
        lengths = [int(v) for v in rng.integers(1, 700, size=n)]
        shards = sweep.shard_balanced(lengths, world)
        assert len(shards) == world
        flat = sorted(int(i) for s in shards for i in s)
        assert flat == list(range(n))
        again = sweep.shard_balanced(lengths, world)
        assert all(np.array_equal(a, b) for a, b in zip(shards, again))
        if n:
            loads = [sum(lengths[int(i)] for i in s) for s in shards]
            assert max(loads) <= sum(lengths) / world + max(lengths)
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= max(1, n)   # no rank starves while work remains
            if n >= world:
                assert min(len(s) for s in shards) >= 1


def test_window_loader_index_logic_matches_the_reference_tables(golden_dir):
    """features.WindowLoader.stage1_windows with the upload and the device gather replaced by their CPU meaning (row t of
    the features holds t, gather = fancy indexing): the windows it asks the device for are the reference's
    (tests/golden/windows.json, produced by the reference's own lines), `--plus_baseline` row included."""
    import json
    from types import SimpleNamespace
    from revisionllm_b200 import features as ft

    class CpuLoader(ft.WindowLoader):
        def __init__(self):
            self.engine = SimpleNamespace(device=torch.device("cpu"), gather_windows=lambda dev, idx: dev[idx.long()])
            self.dim = 1

        def upload(self, features):
            return torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
    g = json.load(open(os.path.join(golden_dir, "windows.json")))
    loader = CpuLoader()
    checked = 0
    for c in g["stage1"]:
        if c["baseline"]:
            continue                                     # --baseline re-samples the features first (:220-222), not a loader mode
        feats = np.arange(c["ctx_l"], dtype=np.float32)[:, None]
        out = loader.stage1_windows(feats, c["debug_window"] * c["feature_fps"], c["num_frames"], plus_baseline=c["plus_baseline"])
        assert out[:, :, 0].long().tolist() == c["windows"], (c["ctx_l"], c["plus_baseline"])
        checked += 1
    assert checked >= 4
    one = loader.stage1_windows(np.arange(77, dtype=np.float32)[:, None], 500, 10, small_video=True)
    assert one[:, :, 0].long().tolist() == [np.linspace(0, 76, 10, dtype=np.int32).tolist()]


class _CpuAdapterEngine:
    """The three entry points the ClipEncoder host code drives (rvl_gemm_bf16, rvl_layernorm, rvl_mha96) restated in torch
    on the CPU with their documented meaning (include/revisionllm_b200.h), so that `clip_encoder.ClipEncoder.__call__` -
    which weights, which biases, which output modes, in which order - can be checked without a GPU."""
    device = torch.device("cpu")

    def gemm(self, A, W, bias=None, out=None, out_mode=_cabi.GEMM_OUT_BF16, flags=0, **kw):
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
            rows = torch.arange(x.shape[0]) % period
            y_pos_bf16.copy_((y + pos.float()[rows]).to(torch.bfloat16))
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
        s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
        if key_mask is not None:
            s = s.masked_fill(key_mask.reshape(-1, Tk)[sel][:, None, None, :] == 0, float("-inf"))
        o = torch.softmax(s, dim=-1) @ vh
        out.copy_(o.permute(0, 2, 1, 3).reshape(n_seq * Tq, n_heads * d).to(torch.bfloat16))


def test_clip_encoder_host_orchestration_matches_reference_fixture(golden_dir):
    """clip_encoder.ClipEncoder.__call__ over a CPU stand-in for the three device entry points, against the reference's
    own adapter output (tests/golden/clip_encoder_tiny.npz): pins the host side of stage 2 - projections, position
    embeddings, global token, post-norm order, shared text K/V through `seg_text_idx`, key mask, CLS projection."""
    from revisionllm_b200.clip_encoder import ClipEncoder
    g = np.load(os.path.join(golden_dir, "clip_encoder_tiny.npz"), allow_pickle=True)
    cw = syn.make_clip_encoder_weights(syn.TINY.hidden, seed=0)
    enc = ClipEncoder(_CpuAdapterEngine(), cw)
    frames, q, qmask = torch.from_numpy(g["frames"]), torch.from_numpy(g["q"]), torch.from_numpy(g["qmask"])
    idx = torch.ones(frames.shape[0], dtype=torch.int32)             # every window attends to text row 1 (masked tail), as in the fixture
    got = enc(frames.to(torch.bfloat16), q.to(torch.bfloat16), qmask, idx).float()
    want = torch.from_numpy(g["cls_out"])
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 3e-2, err                                            # bf16 weights / activations against the fp32 reference


def test_plan_splice_with_several_placeholders_per_row_equals_oracle():
    """Rows with 0, 1, 2 or 3 <video> placeholders: visual blocks are consumed in order, one per placeholder (a row without
    one still consumes a block, vtimellm_arch.py:168-176).  The general (non-vectorised) path of engine.plan_splice against
    oracle/splice_ref.splice on 40 random batches, truncation included."""
    rng = np.random.default_rng(29)
    H, V = 8, 50
    w = {"model.embed_tokens.weight": torch.from_numpy(rng.standard_normal((V, H)).astype(np.float32))}
    for case in range(40):
        B = int(rng.integers(1, 5))
        Ltxt = int(rng.integers(5, 11))
        ids = rng.integers(3, V, size=(B, Ltxt)).astype(np.int64)
        n_blocks = 0
        for b in range(B):
            k = int(rng.integers(0, 4))
            for pos in rng.choice(Ltxt, size=k, replace=False):
                ids[b, int(pos)] = -200
            n_blocks += max(k, 1)
        frames = [int(rng.integers(1, 6)) for _ in range(n_blocks)]
        max_len = None if case % 3 else int(rng.integers(3, Ltxt + 8))
        proj = [torch.from_numpy(rng.standard_normal((f, H)).astype(np.float32)) for f in frames]
        want = splice_ref.splice(w, torch.from_numpy(ids), proj, max_length=max_len)
        plan = plan_splice(ids, frames, None, max_length=max_len)
        assert plan["lengths"].tolist() == [e.shape[0] for e in want], case
        allproj = torch.cat(proj)
        packed = torch.full((int(plan["cu_seqlens"][-1]), H), float("nan"))
        packed[torch.from_numpy(plan["text_dst"]).long()] = w["model.embed_tokens.weight"][torch.from_numpy(plan["text_ids"]).long()]
        packed[torch.from_numpy(plan["vis_dst"]).long()] = allproj[torch.from_numpy(plan["vis_src"]).long()]
        cu = plan["cu_seqlens"]
        for b in range(B):
            assert torch.equal(packed[cu[b]:cu[b + 1]], want[b]), (case, b)


def test_pad_sequences_1d_matches_the_reference_fixture(golden_dir):
    """tests/golden/pad_sequences.npz holds what the reference's own pad_sequences_1d returned (make_golden_pad.py) for ragged
    torch / numpy inputs, nested lists, a fixed length and the stage-1 driver's call (eval_nlq_negative.py:286)."""
    import importlib.util
    from revisionllm_b200.tensor_utils import pad_sequences_1d
    spec = importlib.util.spec_from_file_location("make_golden_pad_cases", os.path.join(golden_dir, "make_golden_pad.py"))
    src = open(spec.origin).read()
    ns = {}
    exec(compile(src[src.index("def cases():"): src.index("def main():")], spec.origin, "exec"), {"torch": torch, "np": np}, ns)
    g = np.load(os.path.join(golden_dir, "pad_sequences.npz"))
    for name, (seqs, kw) in ns["cases"]().items():
        padded, mask = pad_sequences_1d(seqs, **kw)
        assert isinstance(padded, torch.Tensor) == bool(g[name + "/is_torch"]) and isinstance(mask, torch.Tensor) == bool(g[name + "/is_torch"])
        got_p = padded.numpy() if isinstance(padded, torch.Tensor) else padded
        got_m = mask.numpy() if isinstance(mask, torch.Tensor) else mask
        assert got_p.dtype == g[name + "/padded"].dtype and got_m.dtype == np.float32
        np.testing.assert_array_equal(got_p, g[name + "/padded"])
        np.testing.assert_array_equal(got_m, g[name + "/mask"])
    with pytest.raises(AssertionError):
        pad_sequences_1d([torch.zeros(2, 3)], dtype=np.float32)          # mismatched container / dtype: same assertion as the reference


def test_build_stamp_tracks_the_sources():
    """`build()` decides by content, not by file times: the stamp next to the library is the digest of csrc/*, the header
    and the flags (the library reaches the GPU box by copy, where every mtime is new)."""
    from revisionllm_b200 import build as b
    b.build()                                   # compiles only when the stamp and the sources disagree
    assert os.path.exists(b.STAMP) and open(b.STAMP).read().strip() == b.sources_digest()
    assert not b.needs_build()
    stamp = open(b.STAMP).read()
    try:
        open(b.STAMP, "w").write("0" * 64 + "\n")
        assert b.needs_build()                  # a library built from other sources is rebuilt whatever its file time says
    finally:
        open(b.STAMP, "w").write(stamp)
