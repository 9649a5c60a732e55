"""Stage-1 dense sweep of a movie and the stage-2 pass, sharded segment-parallel across ranks.

What the reference does with independent 1-GPU jobs that append JSONL files merged later by file
reads (/root/reference/revisionllm/eval/eval_nlq_negative.py:138,179-180,281-298,337;
/root/reference/revisionllm/eval/metric_retrieval_forward.py:59-79) is done here inside one
torchrun job: segment i of a movie goes to rank i mod R, every rank runs projector + splice +
prefill + greedy decode + cosine scoring on its shard, and ONE small all-gather of a fixed-size
per-segment record puts all results on every rank before stage-2 selection.  There is no other
data-path collective: weights are replicated, segments are independent.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import scoring
from .model import WindowBank

REC_TOKENS = 16
# int32 words: 16 tokens | n_tokens | span_start | span_end | H_mean | H_max | cos   (floats stored as raw bits)
REC_WORDS = REC_TOKENS + 6


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Round-robin: segment i -> rank i mod world."""
    return np.arange(rank, n, world, dtype=np.int64)


def shard_balanced(lengths: Sequence[int], world: int) -> List[np.ndarray]:
    """Ragged batches (VidChapters: windows of 1-100 frames, queries of 8-32 tokens): greedy length-balanced bin packing
    (longest first onto the least loaded rank).  Deterministic - every rank computes the same
    assignment from the same lengths: ties go to the lower index / lower rank.  Returns one ascending index array per rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.lexsort((np.arange(lengths.shape[0]), -lengths))          # by length descending, then index
    load = np.zeros(world, dtype=np.int64)
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))                                          # first minimum = lowest rank on ties
        bins[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.asarray(sorted(b), dtype=np.int64) for b in bins]


def allgather_indexed(local: torch.Tensor, shards: Sequence[np.ndarray], rank: int, world: int, group=None) -> torch.Tensor:
    """All-gather for an arbitrary (but globally known) assignment: rank r holds the records of `shards[r]` in that order;
    every rank receives all records in global index order.  One collective, padded to the largest shard."""
    n_total = int(sum(len(s) for s in shards))
    if world == 1:
        out = torch.empty((n_total, local.shape[1]), dtype=local.dtype, device=local.device)
        out[torch.from_numpy(shards[0]).to(local.device)] = local
        return out
    import torch.distributed as dist
    per = max(len(s) for s in shards)
    buf = torch.full((per, local.shape[1]), -1, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    gathered = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    out = torch.empty((n_total, local.shape[1]), dtype=local.dtype, device=local.device)
    for r, idx in enumerate(shards):
        if len(idx):
            out[torch.from_numpy(idx).to(local.device)] = gathered[r * per: r * per + len(idx)]
    return out


def pack_records(tokens: torch.Tensor, spans: torch.Tensor, h_mean: torch.Tensor, h_max: torch.Tensor,
                 cos: torch.Tensor) -> torch.Tensor:
    """tokens [n, T'<=16] int32, spans [n, 2] int32 (-1 = 'Not Present'), h_mean/h_max/cos [n] fp32 ->
    [n, REC_WORDS] int32."""
    n, t = tokens.shape
    rec = torch.full((n, REC_WORDS), -1, dtype=torch.int32, device=tokens.device)
    tt = min(t, REC_TOKENS)
    rec[:, :tt] = tokens[:, :tt].to(torch.int32)
    rec[:, REC_TOKENS] = tt
    rec[:, REC_TOKENS + 1: REC_TOKENS + 3] = spans.to(torch.int32)
    rec[:, REC_TOKENS + 3] = h_mean.to(torch.float32).contiguous().view(torch.int32)
    rec[:, REC_TOKENS + 4] = h_max.to(torch.float32).contiguous().view(torch.int32)
    rec[:, REC_TOKENS + 5] = cos.to(torch.float32).contiguous().view(torch.int32)
    return rec


def unpack_records(rec: torch.Tensor) -> Dict[str, torch.Tensor]:
    f = lambda c: rec[:, c].contiguous().view(torch.float32)
    return dict(tokens=rec[:, :REC_TOKENS], n_tokens=rec[:, REC_TOKENS], spans=rec[:, REC_TOKENS + 1: REC_TOKENS + 3],
                h_mean=f(REC_TOKENS + 3), h_max=f(REC_TOKENS + 4), cos=f(REC_TOKENS + 5))


def allgather_records(local: torch.Tensor, n_total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """The single exchange step: every rank contributes its round-robin shard, every rank receives all
    `n_total` records in global segment order.  A pure copy, so downstream selection is bit-exact."""
    if world == 1:
        return local
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    buf = torch.full((per, REC_WORDS), -1, dtype=torch.int32, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * per, REC_WORDS), dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    # rank r, slot j  ->  global segment j * world + r
    out = out.view(world, per, REC_WORDS).transpose(0, 1).reshape(world * per, REC_WORDS)
    return out[:n_total].contiguous()


@dataclass
class SweepResult:
    records: torch.Tensor          # [n_segments, REC_WORDS] int32, global order, on every rank
    local_indices: np.ndarray      # segments this rank scored
    stage2_indices: Optional[torch.Tensor] = None


def span_slices(spans: torch.Tensor, n_frames: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row ranges [begin, end) inside each window for the cosine score of its predicted span, as the stage-1 driver takes
    them (/root/reference/revisionllm/eval/eval_nlq_negative.py:79-96,309-310): `features[k][from : to + 1]`, a one-frame
    span widened by one frame on both sides (:90-92), Python slicing clamping at the window's end.  spans [n, 2] int32 with
    -1 for windows whose answer holds no span ("Not Present") - the reference computes no cosine for those; here they score
    over the whole window so that every record carries a value."""
    lo, hi = spans[:, 0].to(torch.int64), spans[:, 1].to(torch.int64)
    valid = (lo >= 0) & (hi >= 0)
    one = valid & (lo == hi)
    lo = torch.where(one, (lo - 1).clamp(min=0), lo)
    hi = torch.where(one, hi + 1, hi)
    begin = torch.where(valid, lo.clamp(max=n_frames), torch.zeros_like(lo))
    end = torch.where(valid, (hi + 1).clamp(max=n_frames), torch.full_like(hi, n_frames))
    return begin, torch.maximum(end, begin)


def score_segments(model, seg_feats: torch.Tensor, input_ids: torch.Tensor, cls: Optional[torch.Tensor],
                   max_new_tokens: int = 16, decode_spans: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                   batch: Optional[int] = None, eos_token_id="config", norm_axis: int = scoring.NORM_ACROSS_FRAMES,
                   ) -> torch.Tensor:
    """Score `seg_feats` [n, F, 768] (host or device): generate up to `max_new_tokens` greedy tokens per
    segment, entropy statistics from the device-side per-step entropies, CLIP cosine top-3 score of each
    segment against `cls` [768].  Returns packed records [n, REC_WORDS] on the model's device.

    The cosine score follows the stage-1 driver (eval_nlq_negative.py:309-316): the frames of the PREDICTED span
    (`decode_spans(new_tokens)` -> [n, 2] frame indices inside the window, -1 = no span; see `span_slices`), normalised
    across the frame axis (`.norm(dim=0)`, the driver's quirk: `norm_axis` = scoring.NORM_ACROSS_FRAMES), sum of the top-3
    similarities.  Without `decode_spans` every window is scored over all its frames."""
    eng = model.engine
    dev = model.device
    n, F, D = seg_feats.shape
    if n == 0:                      # a rank whose shard is empty (fewer windows than ranks) still joins the all-gather
        return torch.empty((0, REC_WORDS), dtype=torch.int32, device=dev)
    batch = n if batch is None else max(1, batch)
    recs = []
    for s in range(0, n, batch):
        feats = seg_feats[s: s + batch].to(dev, torch.bfloat16, non_blocking=True)
        b = feats.shape[0]
        ids = input_ids[None].expand(b, -1) if input_ids.dim() == 1 else input_ids[s: s + b]
        out = model.generate(ids, images=feats, max_new_tokens=max_new_tokens, output_scores=False,
                             return_dict_in_generate=True, eos_token_id=eos_token_id)
        new_tok = out["sequences"][:, ids.shape[1]:].to(torch.int32)
        ent = out["entropies"]
        stats = scoring.entropy_stats_from_steps(ent)
        spans = decode_spans(new_tok).to(dev) if decode_spans is not None else torch.full((b, 2), -1, dtype=torch.int32, device=dev)
        if cls is not None:
            base = torch.arange(0, b * F, F, dtype=torch.int64, device=dev)
            begin, end = span_slices(spans, F)
            cos, _ = eng.cosine_topk(feats.reshape(b * F, D), (base + begin).to(torch.int32), cls.to(dev, torch.bfloat16).contiguous(),
                                     k=3, norm_axis=norm_axis, max_seg_rows=F, want_idx=False, seg_ends=(base + end).to(torch.int32))
        else:
            cos = torch.zeros(b, dtype=torch.float32, device=dev)
        recs.append(pack_records(new_tok, spans, stats[:, 2], stats[:, 0], cos))
    return torch.cat(recs, dim=0)


def score_segments_queries(model, seg_feats: torch.Tensor, input_ids: torch.Tensor, cls: Optional[torch.Tensor],
                           max_new_tokens: int = 16, decode_spans: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                           batch_segments: Optional[int] = None, eos_token_id="config", norm_axis: int = scoring.NORM_ACROSS_FRAMES,
                           attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Several queries over the same segments in one pass - the reference's evaluation loop asks every query of a movie about
    every window of that movie (/root/reference/revisionllm/eval/eval_nlq_negative.py:183-337 runs once per query; MAD has
    hundreds of queries per movie), and SURVEY.md section 8e names batching the queries' sweeps as the way to keep the decode
    batch large when a movie is spread over many GPUs.

    `seg_feats` [S, F, 768]; `input_ids` [Q, Ltxt] (right-padded, `attention_mask` [Q, Ltxt] when the queries differ in
    length); `cls` [Q, 768] or None.  Rows are segment-major - (segment s, query q) is row s * Q + q - so the prompts that
    share a segment sit next to each other: the features travel once (`image_index`), the decode batch is Q times larger for
    the same weight stream, and with `model.share_prefix_compute` the system text and the visual positions of a segment run
    through the decoder once for all Q prompts.  Returns records [S * Q, REC_WORDS] in that row order."""
    eng, dev = model.engine, model.device
    S, F, D = seg_feats.shape
    Q = input_ids.shape[0]
    if S == 0 or Q == 0:
        return torch.empty((0, REC_WORDS), dtype=torch.int32, device=dev)
    step = S if batch_segments is None else max(1, batch_segments)
    recs = []
    for s0 in range(0, S, step):
        feats = seg_feats[s0: s0 + step].to(dev, torch.bfloat16, non_blocking=True)
        b = feats.shape[0]
        ids = input_ids.repeat(b, 1)                                                   # row = segment * Q + query
        am = None if attention_mask is None else attention_mask.repeat(b, 1)
        index = torch.arange(b).repeat_interleave(Q)
        out = model.generate(ids, images=feats, image_index=index, attention_mask=am, max_new_tokens=max_new_tokens, output_scores=False,
                             return_dict_in_generate=True, eos_token_id=eos_token_id)
        new_tok = out["sequences"][:, ids.shape[1]:].to(torch.int32)
        stats = scoring.entropy_stats_from_steps(out["entropies"])
        n = b * Q
        spans = decode_spans(new_tok).to(dev) if decode_spans is not None else torch.full((n, 2), -1, dtype=torch.int32, device=dev)
        cos = torch.zeros(n, dtype=torch.float32, device=dev)
        if cls is not None:
            base = torch.arange(0, b * F, F, dtype=torch.int64, device=dev)
            flat = feats.reshape(b * F, D)
            for q in range(Q):                                                         # one launch per query: its own text vector
                begin, end = span_slices(spans[q::Q], F)
                c, _ = eng.cosine_topk(flat, (base + begin).to(torch.int32), cls[q].to(dev, torch.bfloat16).contiguous(), k=3,
                                       norm_axis=norm_axis, max_seg_rows=F, want_idx=False, seg_ends=(base + end).to(torch.int32))
                cos[q::Q] = c
        recs.append(pack_records(new_tok, spans, stats[:, 2], stats[:, 0], cos))
    return torch.cat(recs, dim=0)


def stage1_sweep_queries(model, segments: torch.Tensor, input_ids: torch.Tensor, cls: Optional[torch.Tensor], max_new_tokens: int = 16,
                         rank: int = 0, world: int = 1, group=None, batch_segments: Optional[int] = None, decode_spans=None,
                         eos_token_id="config", attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`score_segments_queries` on `world` ranks: segment i -> rank i mod world (each rank runs all Q queries on its segments),
    one all-gather.  Returns records [W * Q, REC_WORDS] on every rank, row = segment * Q + query."""
    W, Q = segments.shape[0], input_ids.shape[0]
    mine = shard_indices(W, rank, world)
    local = score_segments_queries(model, segments[torch.from_numpy(mine)], input_ids, cls, max_new_tokens, decode_spans, batch_segments,
                                   eos_token_id, attention_mask=attention_mask)
    if world == 1:
        return local
    # the record gather works on whole segments: Q records of a segment travel together
    wide = allgather_wide(local.view(len(mine), Q * REC_WORDS), W, rank, world, group)
    return wide.view(W * Q, REC_WORDS)


def allgather_wide(local: torch.Tensor, n_total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """`allgather_records` for rows of any width (round-robin shards: rank r, slot j -> global row j * world + r)."""
    import torch.distributed as dist
    width = local.shape[1]
    per = (n_total + world - 1) // world
    buf = torch.full((per, width), -1, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * per, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out.view(world, per, width).transpose(0, 1).reshape(world * per, width)[:n_total].contiguous()


def stage1_sweep(model, segments: torch.Tensor, input_ids: torch.Tensor, cls: Optional[torch.Tensor],
                 max_new_tokens: int = 16, rank: int = 0, world: int = 1, group=None, batch: Optional[int] = None,
                 decode_spans=None, eos_token_id="config", stage2_topk: Optional[int] = None,
                 norm_axis: int = scoring.NORM_ACROSS_FRAMES) -> SweepResult:
    """`segments` [W, F, 768]: all windows of the movie (every rank holds the same host tensor; only the
    local shard is copied to the GPU)."""
    W = segments.shape[0]
    mine = shard_indices(W, rank, world)
    local = score_segments(model, segments[torch.from_numpy(mine)], input_ids, cls, max_new_tokens, decode_spans, batch,
                           eos_token_id, norm_axis)
    allrec = allgather_records(local, W, rank, world, group)
    res = SweepResult(allrec, mine)
    if stage2_topk is not None:
        cos = unpack_records(allrec)["cos"]
        res.stage2_indices = scoring.select_topk_segments(model.engine, cos, stage2_topk)
    return res


def ragged_sweep(model, windows: Sequence[torch.Tensor], input_ids: torch.Tensor, attention_mask: torch.Tensor,
                 cls: Optional[torch.Tensor], max_new_tokens: int = 16, rank: int = 0, world: int = 1, group=None,
                 max_tokens_per_batch: int = 1 << 16, eos_token_id="config") -> torch.Tensor:
    """BASELINE.json config 5 (VidChapters-shaped): windows with different frame counts `windows[i]` [F_i, 768] and
    right-padded prompts `input_ids` [n, Ltxt] / `attention_mask` of different lengths, varlen-packed.  Windows are dealt
    to the ranks by `shard_balanced` on their spliced lengths, each rank scores its share in packed batches of at most
    `max_tokens_per_batch` prompt tokens, and one `allgather_indexed` returns all records in window order."""
    n = len(windows)
    am = attention_mask.bool()
    lengths = [int(am[i].sum()) - 1 + int(windows[i].shape[0]) for i in range(n)]
    shards = shard_balanced(lengths, world)
    mine = shards[rank]
    eng, dev = model.engine, model.device
    recs = []
    start = 0
    while start < len(mine):
        tok, end = 0, start
        while end < len(mine) and (end == start or tok + lengths[mine[end]] <= max_tokens_per_batch):
            tok += lengths[mine[end]]
            end += 1
        sel = [int(i) for i in mine[start:end]]
        imgs = [windows[i] for i in sel]
        out = model.generate(input_ids[sel], images=imgs, attention_mask=am[sel], max_new_tokens=max_new_tokens, output_scores=False,
                             return_dict_in_generate=True, eos_token_id=eos_token_id)
        new_tok = out["sequences"][:, input_ids.shape[1]:].to(torch.int32)
        stats = scoring.entropy_stats_from_steps(out["entropies"])
        if cls is not None:
            rows = torch.cat([w.reshape(-1, w.shape[-1]) for w in imgs]).to(dev, torch.bfloat16).contiguous()
            offs = torch.tensor(np.concatenate([[0], np.cumsum([w.shape[0] for w in imgs])]), dtype=torch.int32, device=dev)
            cos, _ = eng.cosine_topk(rows, offs, cls.to(dev, torch.bfloat16).contiguous(), k=3, max_seg_rows=max(w.shape[0] for w in imgs))
        else:
            cos = torch.zeros(len(sel), dtype=torch.float32, device=dev)
        spans = torch.full((len(sel), 2), -1, dtype=torch.int32, device=dev)
        recs.append(pack_records(new_tok, spans, stats[:, 2], stats[:, 0], cos))
        start = end
    local = torch.cat(recs, dim=0) if recs else torch.empty((0, REC_WORDS), dtype=torch.int32, device=dev)
    return allgather_indexed(local, shards, rank, world, group)


def _stage2_plan(N: int, batch: int, zooms: Sequence[int], gen: Optional[torch.Generator]) -> List[Dict]:
    """The generate() calls of one query's hierarchical pass, in the reference's order (e2e2:337-353): for each zoom,
    chunks of batch // zoom windows (the last chunk is shifted back to full size), permuted, each window repeated `zoom` times.
    With fewer windows than one chunk the shifted start goes negative and the reference's `features[start:end]` follows
    Python's slicing rules (20 windows, chunks of 25: start = -5 keeps the LAST five; 12 windows: start = -13 keeps all);
    `start` is kept as computed because the answer is mapped back through `starts[i] + j` (:373).  Pinned by
    tests/golden/stage2_plan.json, which the reference's own loop produced."""
    calls: List[Dict] = []
    ids = torch.arange(N)
    for zoom in zooms:
        b = max(1, batch // zoom)
        n_chunks = (N + b - 1) // b
        for i in range(n_chunks):
            start = i * b
            end = min(start + b, N)
            if end - start < b:
                start = end - b
            chunk = ids[start:end]
            n = int(chunk.shape[0])
            idx = torch.randperm(n, generator=gen) if gen is not None else torch.arange(n)
            rows = chunk[idx].repeat_interleave(zoom) if zoom > 1 else chunk[idx]
            calls.append(dict(zoom=zoom, start=start, n=n, idx=idx, rows=rows))
    return calls


def _stage2_finish(calls: List[Dict], results: List[Dict], grounding_windows: Sequence[int], answer_number) -> List[Dict]:
    out: List[Dict] = []
    for c, r in zip(calls, results):
        new_tok, stats = r["tokens"], r["stats"]
        zoom, start, n, idx = c["zoom"], c["start"], c["n"], c["idx"]
        number = answer_number(new_tok) if answer_number is not None else None
        picked = None
        if number is not None:
            j = number // zoom
            if j < n:
                j = int(idx[j])
            j = min(max(start + j, 0), len(grounding_windows) - 1)
            picked = int(grounding_windows[j])
        out.append(dict(zoom=zoom, start=start, perm=idx.tolist(), tokens=new_tok.tolist(), inv_max_entropy=1.0 / float(stats[0]),
                        inv_mean_entropy=1.0 / float(stats[2]), window=picked))
    return out


def stage2_pass(model, windows: torch.Tensor, query_feats, input_ids: torch.Tensor, grounding_windows: Sequence[int],
                batch: int = 100, zooms: Sequence[int] = (4, 2, 1), max_new_tokens: int = 16, perm_seed: Optional[int] = 0,
                answer_number: Optional[Callable[[torch.Tensor], Optional[int]]] = None, eos_token_id="config",
                max_calls_per_batch: int = 16, dedup: bool = True, rank: int = 0, world: int = 1, shard_calls: bool = False,
                group=None) -> List[Dict]:
    """Stage-2 hierarchical pass (/root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:337-386).

    `windows` [N, T, 768]: the selected stage-2 windows (already restricted to `grounding_windows`).  For each
    zoom in (4, 2, 1): chunks of batch // zoom windows, permuted, each repeated `zoom` times, go through ONE
    generate() call as a [1, batch, T, 768] hierarchy input (one ClipEncoder CLS token per window); the first
    integer of the answer // zoom indexes the permuted chunk and is mapped back to a window id.
    The reference permutes with an unseeded torch.randperm (:348); here the permutation comes from
    `perm_seed` (None = identity) so runs are reproducible.
    All chunks of all zoom levels are independent prompts, so they run as one batched generate() (up to
    `max_calls_per_batch` rows; 1 restores the reference's one-call-per-chunk schedule), and the adapter sees every window
    once (`dedup`, see stage2_pass_queries; False stacks the zoom repeats like the reference).
    Returns one dict per chunk (the reference's generate() calls, in its order): tokens, entropy stats (1/max, 1/mean as in
    :356-359), picked window."""
    q = dict(windows=windows, query_feats=query_feats, input_ids=input_ids, grounding_windows=grounding_windows, perm_seed=perm_seed)
    if not shard_calls:
        rank, world = 0, 1                                   # one query: it runs on the calling rank
    return stage2_pass_queries(model, [q], batch, zooms, max_new_tokens, answer_number, eos_token_id, max_calls_per_batch, rank, world,
                               dedup=dedup, shard_calls=shard_calls, group=group)[0]


def stage2_pass_queries(model, queries: Sequence[Dict], batch: int = 100, zooms: Sequence[int] = (4, 2, 1), max_new_tokens: int = 16,
                        answer_number: Optional[Callable[[torch.Tensor], Optional[int]]] = None, eos_token_id="config",
                        max_calls_per_batch: int = 64, rank: int = 0, world: int = 1, dedup: bool = True,
                        shard_calls: bool = False, group=None) -> List[Optional[List[Dict]]]:
    """Stage 2 for SEVERAL queries at once (north star: "one GPU per query, batched across queries").

    Each entry of `queries` holds what `stage2_pass` takes for one query: `windows` [N, T, 768], `query_feats`
    ((tokens [1, Lq, 768], mask [1, Lq]) or None), `input_ids` [L], `grounding_windows`, optional `perm_seed`.
    Query i belongs to rank i mod world (no exchange step: a query's result stays on its rank; the other entries of the
    returned list are None).  Every chunk of every zoom level of every local query is an independent prompt, so they are
    batched across queries into generate() calls of up to `max_calls_per_batch` rows: the 13 GB of weights are streamed once
    per decode step for all of them instead of once per chunk.  Prompts of different lengths are right-padded and masked;
    query features of different lengths are padded with masked rows.  Per query the result equals `stage2_pass`'s.
    `dedup` (default): the ClipEncoder adapter sees every distinct (query, window) pair once instead of once per zoom repeat
    and per chunk (7 x fewer adapter rows for zooms (4, 2, 1)); False feeds the stacked copies like the reference.
    `shard_calls` (fewer queries than ranks - one movie-query on 8 GPUs): every rank holds every query, the generate() calls
    of all queries are dealt round-robin to the ranks instead, and one all-gather of fixed-size records (the stage-1 record
    layout: tokens, entropy mean / max) gives every rank every call's result; needs max_new_tokens <= REC_TOKENS."""
    shard_calls = bool(shard_calls) and world > 1 and max_new_tokens <= REC_TOKENS
    mine = [qi for qi in range(len(queries)) if shard_calls or qi % world == rank]
    jobs: List[Tuple[int, int]] = []                       # (query, call)
    plans: Dict[int, List[Dict]] = {}
    for qi in mine:
        q = queries[qi]
        seed = q.get("perm_seed", 0)
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        plans[qi] = _stage2_plan(int(q["windows"].shape[0]), batch, zooms, gen)
        jobs += [(qi, ci) for ci in range(len(plans[qi]))]
    all_jobs = jobs
    if shard_calls:
        jobs = all_jobs[rank::world]
    results: Dict[Tuple[int, int], Dict] = {}
    groups: Dict[Tuple[int, int, bool], List[Tuple[int, int]]] = {}
    for (qi, ci) in jobs:                                   # one generate() needs equal visual rows / frames per row
        w = queries[qi]["windows"]
        has_q = queries[qi].get("query_feats") is not None
        # the window bank takes prompts with different numbers of visual rows in one batch; stacked 4-D images need them equal
        key = (0 if (has_q and dedup) else int(plans[qi][ci]["rows"].shape[0]), int(w.shape[1]), has_q)
        groups.setdefault(key, []).append((qi, ci))
    for (v_rows, _, has_q), members in groups.items():
        for g0 in range(0, len(members), max_calls_per_batch):
            part = members[g0: g0 + max_calls_per_batch]
            L = max(int(queries[qi]["input_ids"].shape[0]) for (qi, _) in part)
            ids = torch.zeros((len(part), L), dtype=torch.int64)
            am = torch.zeros((len(part), L), dtype=torch.bool)
            for r, (qi, _) in enumerate(part):
                row = queries[qi]["input_ids"].cpu()
                ids[r, : row.shape[0]] = row
                am[r, : row.shape[0]] = True
            part_q = sorted({qi for (qi, _) in part})
            qf = None
            if has_q:
                Lq = max(int(queries[qi]["query_feats"][0].shape[1]) for qi in part_q)
                t0, m0 = queries[part_q[0]]["query_feats"]
                qt = torch.zeros((len(part_q), Lq, t0.shape[-1]), dtype=t0.dtype, device=t0.device)
                qm = torch.zeros((len(part_q), Lq), dtype=m0.dtype, device=m0.device)
                for r, qi in enumerate(part_q):
                    t, m = queries[qi]["query_feats"]
                    qt[r, : t.shape[1]] = t[0]
                    qm[r, : m.shape[1]] = m[0]
                qf = (qt, qm)
            if has_q and dedup:
                # every distinct (query, window) of this batch goes through the adapter once; zoom repeats and the chunks of
                # the three zoom levels that show the same window reuse its CLS row (model.WindowBank)
                base, banks, text_index = {}, [], []
                for r, qi in enumerate(part_q):
                    used = torch.unique(torch.cat([plans[qi][ci]["rows"] for (qj, ci) in part if qj == qi]))
                    lut = torch.full((int(queries[qi]["windows"].shape[0]),), -1, dtype=torch.int64)
                    lut[used] = torch.arange(used.shape[0]) + sum(int(b.shape[0]) for b in banks)
                    base[qi] = lut
                    banks.append(queries[qi]["windows"][used.to(queries[qi]["windows"].device)])
                    text_index += [r] * int(used.shape[0])
                rows = [base[qi][plans[qi][ci]["rows"]] for (qi, ci) in part]
                feat = WindowBank(torch.cat(banks), torch.stack(rows) if len({int(r.shape[0]) for r in rows}) == 1 else rows,
                                  torch.tensor(text_index, dtype=torch.int32))
            else:
                feat = torch.stack([queries[qi]["windows"][plans[qi][ci]["rows"].to(queries[qi]["windows"].device)] for (qi, ci) in part])
                if has_q:                                   # one text row per prompt, as generate() expects for 4-D images
                    pos = {qi: r for r, qi in enumerate(part_q)}
                    sel = torch.tensor([pos[qi] for (qi, _) in part])
                    qf = (qf[0][sel.to(qf[0].device)], qf[1][sel.to(qf[1].device)])
            res = model.generate(ids, images=feat, query_feats=qf, attention_mask=None if bool(am.all()) else am,
                                 max_new_tokens=max_new_tokens, output_scores=False, return_dict_in_generate=True, eos_token_id=eos_token_id,
                                 mask_entropy_after_eos=True)      # a prompt's statistics stop at its own EOS, as in the reference's one call per chunk (:353-359)
            stats_all = scoring.entropy_stats_from_steps(res["entropies"])
            for r, job in enumerate(part):
                results[job] = dict(tokens=res["sequences"][r, L:], stats=stats_all[r])
    if shard_calls:
        dev = model.device
        n_loc = len(jobs)
        tok = torch.full((n_loc, REC_TOKENS), -1, dtype=torch.int32, device=dev)
        st = torch.zeros((n_loc, 4), dtype=torch.float32, device=dev)
        width = torch.zeros(n_loc, dtype=torch.int32, device=dev)
        for i, job in enumerate(jobs):
            t = results[job]["tokens"]
            tok[i, : t.shape[0]] = t.to(dev, torch.int32)
            width[i] = int(t.shape[0])
            st[i] = torch.as_tensor(results[job]["stats"], dtype=torch.float32).to(dev)
        local = pack_records(tok, torch.full((n_loc, 2), -1, dtype=torch.int32, device=dev), st[:, 2], st[:, 0], st[:, 1])
        local[:, REC_TOKENS] = width                        # pack_records stores the buffer width; a call may have produced fewer tokens
        shards = [np.arange(r, len(all_jobs), world, dtype=np.int64) for r in range(world)]
        un = unpack_records(allgather_indexed(local, shards, rank, world, group).cpu())
        results = {}
        for i, job in enumerate(all_jobs):
            n_t = int(un["n_tokens"][i])
            # stats in entropy_stats_from_steps order (max, min, mean, std): the consumers read max and mean
            stats = torch.stack([un["h_max"][i], un["cos"][i], un["h_mean"][i], torch.zeros(())])
            results[job] = dict(tokens=un["tokens"][i, :n_t].to(torch.int64), stats=stats)
    out: List[Optional[List[Dict]]] = [None] * len(queries)
    for qi in mine:
        out[qi] = _stage2_finish(plans[qi], [results[(qi, ci)] for ci in range(len(plans[qi]))], queries[qi]["grounding_windows"], answer_number)
    return out


# ------------------------------------------------------------------------------------------ one movie, end to end
@dataclass
class MovieConfig:
    """Geometry and scoring switches of the two evaluation scripts (their argparse defaults)."""
    clip_length: int = 250            # stage-1 window length in feature frames (debug_window * feature_fps)
    num_frames: int = 100             # frames sampled per stage-1 window
    stage2_clip_length: int = 250     # stage-2 window length in feature frames
    stage2_num_frames: int = 250      # frames per stage-2 window (through the ClipEncoder)
    stride: int = 5                   # stage-2 windows start every clip_length // stride frames
    batch: int = 100                  # stage-2 windows per query ("top-100"; 33 for stage2_long_33)
    zooms: Tuple[int, ...] = (4, 2, 1)
    max_new_tokens: int = 16
    score_merge: str = "multiply"     # eval_nlq_negative.py --score_merge
    normalize: bool = True
    perm_seed: Optional[int] = 0
    stage1_batch: Optional[int] = None
    stage2_calls_per_batch: int = 64  # stage-2 chunks (all zoom levels) that share one batched generate()
    stage2_shard_calls: bool = True   # world > 1: deal the stage-2 generate() calls of the query to all ranks (one more all-gather)


@dataclass
class MovieResult:
    records: torch.Tensor                       # stage-1 records of every window, global order (on every rank)
    answers: List[str]                          # stage-1 answer per window
    clip_frames: Dict[int, Tuple[int, int]]     # window -> predicted span (windows whose answer names one)
    ious: List[float]
    grounding_windows: List[int]                # stage-2 windows chosen from the stage-1 answers
    stage2: Optional[List[Dict]]                # one entry per stage-2 generate() call (None on ranks that did not run it)
    stage2_answers: Optional[List[str]]
    stage2_frames: Optional[Dict[int, Tuple[int, int]]]
    ranked: Optional[Dict[str, list]]           # proposals best first: windows / scores / ious


def run_movie(model, features, input_ids: torch.Tensor, cls: torch.Tensor, detok: Callable[[torch.Tensor], List[str]],
              gt: Tuple[float, float], cfg: MovieConfig = MovieConfig(), query_feats=None, stage2_input_ids: Optional[torch.Tensor] = None,
              detok_stage2: Optional[Callable[[torch.Tensor], List[str]]] = None, rank: int = 0, world: int = 1, group=None,
              stage2_rank: int = 0, eos_token_id="config", timings: Optional[Dict[str, float]] = None) -> MovieResult:
    """One movie-query through the whole recursive path, as ONE call on `world` GPUs - what the reference spreads over three
    scripts and their JSONL files:

      stage 1  (/root/reference/revisionllm/eval/eval_nlq_negative.py:221-336)   half-overlapping windows -> generate ->
               answers + entropy statistics -> predicted spans -> cosine score of each span; windows dealt round-robin to the
               ranks, one all-gather of the fixed-size records;
      select   (/root/reference/revisionllm/eval/eval_nlq_retrieval_e2e2.py:262-294)  stage-2 window grid, windows of the
               positive stage-1 answers padded evenly to `batch`;
      stage 2  (:337-386)  zoom levels 4 / 2 / 1 over the chosen windows through the ClipEncoder: the independent generate()
               calls of the query dealt to the ranks + one all-gather of their records (`cfg.stage2_shard_calls`), or all on
               `stage2_rank` (north star: one GPU per query);
      rank     (/root/reference/revisionllm/eval/metric_retrieval_forward.py:96-199)  stage-1 proposals kept where stage 2
               looked, merged cosine / entropy score, best first (`rvl_merge_rank`).

    `features` [T, 768] fp32 (host numpy or torch); `detok(tokens [n, T'])` -> answer strings (a tokenizer's batch_decode with
    the reference's strip / stop-string rule); `gt` = (start, end) as fractions of the movie.  Every rank returns the stage-1
    part; the ranking lives on `stage2_rank`, the stage-2 calls on every rank (on `stage2_rank` only without call sharding).
    `timings`: a dict that receives the wall-clock ms of each phase (the device is synchronised at the phase boundaries: a
    diagnostic, it serialises what otherwise overlaps)."""
    import time
    from . import metrics
    from .features import WindowLoader
    eng, dev = model.engine, model.device
    t_prev = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + 1e3 * (now - t_prev[0])
            t_prev[0] = now
    feats_np = features.numpy() if isinstance(features, torch.Tensor) else np.asarray(features)
    T = feats_np.shape[0]
    loader = getattr(model, "_window_loader", None)                                        # pinned staging buffer, kept across movies
    if loader is None or loader.dim != feats_np.shape[1]:
        loader = model._window_loader = WindowLoader(eng, max_frames=max(T, 1 << 15), dim=feats_np.shape[1])
    movie = loader.upload(feats_np)                                                        # fp32 [T, 768] on the device, once
    idx1 = scoring.stage1_windows(T, cfg.clip_length, cfg.num_frames)
    if idx1.shape[0] == 0:                                                                 # shorter than one window: one uniform sample
        idx1 = np.linspace(0, T - 1, cfg.num_frames, dtype=np.int32)[None]
    W = idx1.shape[0]
    mine = shard_indices(W, rank, world)
    win1 = eng.gather_windows(movie, torch.from_numpy(np.ascontiguousarray(idx1[mine])).to(dev)) if len(mine) else \
        torch.empty((0, cfg.num_frames, feats_np.shape[1]), dtype=torch.bfloat16, device=dev)
    lap("upload_and_window_gather")

    def decode_spans(tok):
        spans = torch.full((tok.shape[0], 2), -1, dtype=torch.int32)
        for i, text in enumerate(detok(tok.cpu())):
            sp = scoring.parse_span(text)
            if sp is not None:
                spans[i, 0], spans[i, 1] = sp
        return spans
    local = score_segments(model, win1, input_ids, cls, cfg.max_new_tokens, decode_spans, cfg.stage1_batch, eos_token_id)
    lap("stage1")
    records = allgather_records(local, W, rank, world, group)
    un = unpack_records(records.cpu())
    answers = detok(un["tokens"][:, : int(un["n_tokens"].max())])
    num_frames_video = int(T * cfg.num_frames / cfg.clip_length)
    clip_frames, ious, ent = metrics.iou(answers, gt, cfg.num_frames, num_frames_video, un["h_mean"].tolist())
    cos = [float(un["cos"][w]) for w in clip_frames]
    # ---- stage-2 selection (every rank computes the same list; only stage2_rank uses it)
    idx2, _ = scoring.stage2_windows(T, cfg.stage2_clip_length, cfg.stage2_num_frames, cfg.stride)
    grounding = scoring.stage2_select_windows(answers, idx2.shape[0], cfg.batch, cfg.stride) if idx2.shape[0] else []
    # the mapping of :281-283 assumes both stages cut windows of the same length; with other geometries it can name windows past
    # the end of the stage-2 grid (the reference then fails on `clip_feats[i]` and its bare `except` drops the query): leave them out
    grounding = [w for w in grounding if -idx2.shape[0] <= w < idx2.shape[0]]
    res = MovieResult(records, answers, clip_frames, ious, grounding, None, None, None, None)
    lap("allgather_parse_select")
    # one query has up to ~24 independent stage-2 prompts (chunks x zoom levels): with several ranks they are dealt out and the
    # per-call records all-gathered, so every rank holds the stage-2 calls and `stage2_rank` ranks them
    sharded = cfg.stage2_shard_calls and world > 1 and cfg.max_new_tokens <= REC_TOKENS
    if (rank != stage2_rank and not sharded) or not grounding or model.clip_encoder is None:
        return res
    gw = np.asarray(grounding, dtype=np.int64)                                             # negative ids index from the end, as in the reference
    win2 = eng.gather_windows(movie, torch.from_numpy(np.ascontiguousarray(idx2[gw])).to(dev))
    dt2 = detok_stage2 or detok

    def answer_number(tok):
        return scoring.parse_first_int(dt2(tok[None].cpu())[0])
    ids2 = stage2_input_ids if stage2_input_ids is not None else input_ids
    lap("stage2_window_gather")
    calls = stage2_pass(model, win2, query_feats, ids2, grounding, cfg.batch, cfg.zooms, cfg.max_new_tokens, cfg.perm_seed,
                        answer_number, eos_token_id, max_calls_per_batch=cfg.stage2_calls_per_batch, rank=rank, world=world,
                        shard_calls=sharded, group=group)
    lap("stage2_pass")
    answers2 = [dt2(torch.tensor(c["tokens"])[None])[0] for c in calls]
    frames2, _hit = metrics.stage2_frames(answers2, gt, cfg.batch, [c["start"] for c in calls], [c["perm"] for c in calls],
                                          [c["zoom"] for c in calls], grounding)
    res.stage2, res.stage2_answers, res.stage2_frames = calls, answers2, frames2
    if clip_frames and rank == stage2_rank:
        # rank_query expects one entry per answered window in window order - `present` there is defined by the answer strings
        present = [i for i, a in enumerate(answers) if a != "Not Present" and a != "From 249 to 249."]
        if present == list(clip_frames):
            res.ranked = metrics.rank_query(eng, answers, cos, ent, ious, stage2_frames=frames2, mode=cfg.score_merge, normalize=cfg.normalize)
    lap("merge_rank")
    return res


def run_movie_queries(model, features, input_ids: torch.Tensor, cls: torch.Tensor, detok: Callable[[torch.Tensor], List[str]],
                      gts: Sequence[Tuple[float, float]], cfg: MovieConfig = MovieConfig(), query_feats: Optional[Sequence] = None,
                      stage2_input_ids: Optional[torch.Tensor] = None, detok_stage2: Optional[Callable[[torch.Tensor], List[str]]] = None,
                      rank: int = 0, world: int = 1, group=None, eos_token_id="config",
                      timings: Optional[Dict[str, float]] = None) -> List[MovieResult]:
    """`run_movie` for Q queries on the SAME movie in one call - the reference runs its three scripts once per query.

      stage 1  every window with every query in one pass (`score_segments_queries`: features stored once per window, with
               `model.share_prefix_compute` the system text + visual positions of a window computed once for its Q prompts);
               windows dealt round-robin to the ranks, one all-gather (the Q records of a window travel together);
      select   per query, on every rank, from the gathered records;
      stage 2  whole queries dealt to the ranks and batched across the queries of a rank (`stage2_pass_queries`; north star: one
               GPU per query, batched across queries) - with fewer queries than ranks the generate() calls are dealt instead;
      rank     per query, on the rank that ran its stage 2.

    `input_ids` [Q, Ltxt] (same length; the system text in front of <video> is common), `cls` [Q, 768], `gts[q]` = (start, end)
    fractions, `query_feats[q]` = (tokens [1, Lq, 768], mask [1, Lq]), `stage2_input_ids` [Q, L2] or [L2].  Returns one
    MovieResult per query; `stage2` / `ranked` are filled on the rank that owns the query (query q -> rank q mod world), on every
    rank when the calls were dealt."""
    import time
    from . import metrics
    from .features import WindowLoader
    eng, dev = model.engine, model.device
    t_prev = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + 1e3 * (now - t_prev[0])
            t_prev[0] = now
    Q = int(input_ids.shape[0])
    feats_np = features.numpy() if isinstance(features, torch.Tensor) else np.asarray(features)
    T = feats_np.shape[0]
    loader = getattr(model, "_window_loader", None)
    if loader is None or loader.dim != feats_np.shape[1]:
        loader = model._window_loader = WindowLoader(eng, max_frames=max(T, 1 << 15), dim=feats_np.shape[1])
    movie = loader.upload(feats_np)
    idx1 = scoring.stage1_windows(T, cfg.clip_length, cfg.num_frames)
    if idx1.shape[0] == 0:
        idx1 = np.linspace(0, T - 1, cfg.num_frames, dtype=np.int32)[None]
    W = idx1.shape[0]
    mine = shard_indices(W, rank, world)
    win1 = eng.gather_windows(movie, torch.from_numpy(np.ascontiguousarray(idx1[mine])).to(dev)) if len(mine) else \
        torch.empty((0, cfg.num_frames, feats_np.shape[1]), dtype=torch.bfloat16, device=dev)
    lap("upload_and_window_gather")

    def decode_spans(tok):
        spans = torch.full((tok.shape[0], 2), -1, dtype=torch.int32)
        for i, text in enumerate(detok(tok.cpu())):
            sp = scoring.parse_span(text)
            if sp is not None:
                spans[i, 0], spans[i, 1] = sp
        return spans
    local = score_segments_queries(model, win1, input_ids, cls, cfg.max_new_tokens, decode_spans, cfg.stage1_batch, eos_token_id)
    lap("stage1")
    records = local if world == 1 else allgather_wide(local.view(len(mine), Q * REC_WORDS), W, rank, world, group).view(W * Q, REC_WORDS)
    rec_cpu = records.cpu()
    num_frames_video = int(T * cfg.num_frames / cfg.clip_length)
    idx2, _ = scoring.stage2_windows(T, cfg.stage2_clip_length, cfg.stage2_num_frames, cfg.stride)
    results: List[MovieResult] = []
    per_q = []
    for q in range(Q):
        un = unpack_records(rec_cpu[q::Q])
        answers = detok(un["tokens"][:, : int(un["n_tokens"].max())])
        clip_frames, ious, ent = metrics.iou(answers, gts[q], cfg.num_frames, num_frames_video, un["h_mean"].tolist())
        cos = [float(un["cos"][w]) for w in clip_frames]
        grounding = scoring.stage2_select_windows(answers, idx2.shape[0], cfg.batch, cfg.stride) if idx2.shape[0] else []
        grounding = [w for w in grounding if -idx2.shape[0] <= w < idx2.shape[0]]
        results.append(MovieResult(records[q::Q], answers, clip_frames, ious, grounding, None, None, None, None))
        per_q.append((cos, ent))
    lap("allgather_parse_select")
    if model.clip_encoder is None or query_feats is None:
        return results
    live = [q for q in range(Q) if results[q].grounding_windows]
    deal_calls = cfg.stage2_shard_calls and world > 1 and len(live) < world and cfg.max_new_tokens <= REC_TOKENS
    dt2 = detok_stage2 or detok
    queries = []
    for q in live:
        gw = np.asarray(results[q].grounding_windows, dtype=np.int64)
        owned = deal_calls or (len(queries) % world == rank)                    # only the owner gathers the query's windows
        win2 = eng.gather_windows(movie, torch.from_numpy(np.ascontiguousarray(idx2[gw])).to(dev)) if owned else None
        ids2 = input_ids[q] if stage2_input_ids is None else (stage2_input_ids if stage2_input_ids.dim() == 1 else stage2_input_ids[q])
        queries.append(dict(windows=win2, query_feats=query_feats[q], input_ids=ids2, grounding_windows=results[q].grounding_windows,
                            perm_seed=cfg.perm_seed))
    lap("stage2_window_gather")
    answer_number = lambda tok: scoring.parse_first_int(dt2(tok[None].cpu())[0])
    passes = stage2_pass_queries(model, queries, cfg.batch, cfg.zooms, cfg.max_new_tokens, answer_number, eos_token_id,
                                 cfg.stage2_calls_per_batch, rank, world, shard_calls=deal_calls, group=group)
    lap("stage2_pass")
    for n, q in enumerate(live):
        calls = passes[n]
        if calls is None:
            continue
        res = results[q]
        answers2 = [dt2(torch.tensor(c["tokens"])[None])[0] for c in calls]
        frames2, _hit = metrics.stage2_frames(answers2, gts[q], cfg.batch, [c["start"] for c in calls], [c["perm"] for c in calls],
                                              [c["zoom"] for c in calls], res.grounding_windows)
        res.stage2, res.stage2_answers, res.stage2_frames = calls, answers2, frames2
        if res.clip_frames and (not deal_calls or rank == 0):
            present = [i for i, a in enumerate(res.answers) if a != "Not Present" and a != "From 249 to 249."]
            if present == list(res.clip_frames):
                cos, ent = per_q[q]
                res.ranked = metrics.rank_query(eng, res.answers, cos, ent, res.ious, stage2_frames=frames2, mode=cfg.score_merge,
                                                normalize=cfg.normalize)
    lap("merge_rank")
    return results
