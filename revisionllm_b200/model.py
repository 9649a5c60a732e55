"""Host-side mirror of the reference model surface, backed by the sm_100a kernels.

Mirrors `VTimeLLMLlamaForCausalLM` (/root/reference/revisionllm/model/vtimellm_llama.py:23-110):
`forward(...)` with the reference's keyword list (:38-56), `generate(input_ids, images=, query_feats=, ...)`
as called by `inference()` (/root/reference/revisionllm/inference.py:45-59), `get_model()`,
`get_input_embeddings()`, `.config`, `.bfloat16().cuda()`.  The splice index math follows
`prepare_inputs_labels_for_multimodal` (/root/reference/revisionllm/model/vtimellm_arch.py:81-299).

Differences, all deliberate and documented in DESIGN.md:
  * decoding is greedy by default (BASELINE.json north_star); `do_sample=True, temperature=T` draws from
    softmax(logits / T) like the reference's `sample()` (vtimellm_llama.py:312-338), with a counter-based Philox stream
    instead of torch.multinomial's;
  * ragged batches are packed (cu_seqlens) instead of right-padded; outputs are re-padded on return;
  * there is no CPU path: `.float()` / CPU placement raise.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import constants
from ._cabi import RvlError
from .engine import Engine, EngineConfig, plan_splice


@dataclass
class WindowBank:
    """Stage-2 hierarchy input without the repetition: the reference stacks every window once per zoom repeat and per chunk
    into `images [b, v, t, 768]` (eval_nlq_retrieval_e2e2.py:348-352) and runs the ClipEncoder adapter over all copies.  The
    CLS row of a window depends only on (window, query), so here `windows [U, t, 768]` holds each distinct pair once,
    `text_index [U]` names the row of `query_feats` it attends to, and `rows [b, v]` (or a list of b index vectors when the
    prompts hold different numbers of windows) says which of the U CLS rows fills each visual position of each prompt.
    Pass it as `images=` to generate() / forward()."""
    windows: torch.Tensor
    rows: object
    text_index: torch.Tensor


@dataclass
class RevisionConfig:
    """Subset of the HF/VTimeLLM config the reference touches (vtimellm_arch.py:106,172,241,257,291)."""
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    vocab_size: int = 32000
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096
    adapter_input_dim: int = 768
    model_type: str = "VTimeLLM"
    tokenizer_model_max_length: Optional[int] = None
    tokenizer_padding_side: str = "right"
    bos_token_id: int = 1
    eos_token_id: int = 2
    pad_token_id: Optional[int] = None
    # stage-2 adapter switches (vtimellm_arch.py:12-31)
    clip_adapter: bool = False
    clip_adapter_text: bool = False
    hierarchy: bool = False

    @classmethod
    def from_synth(cls, s, **kw) -> "RevisionConfig":
        return cls(hidden_size=s.hidden, intermediate_size=s.intermediate, num_hidden_layers=s.n_layers,
                   num_attention_heads=s.n_heads, vocab_size=s.vocab, rms_norm_eps=s.rms_eps, rope_theta=s.rope_theta,
                   max_position_embeddings=s.max_pos, adapter_input_dim=s.adapter_dim, **kw)

    def engine_config(self) -> EngineConfig:
        return EngineConfig(hidden=self.hidden_size, n_layers=self.num_hidden_layers, n_heads=self.num_attention_heads,
                            head_dim=self.hidden_size // self.num_attention_heads, intermediate=self.intermediate_size,
                            vocab=self.vocab_size, adapter_dim=self.adapter_input_dim, max_pos=self.max_position_embeddings,
                            rms_eps=self.rms_norm_eps, rope_theta=self.rope_theta)


class KVState:
    """What `past_key_values` is here: the paged-cache bookkeeping of one live batch."""

    def __init__(self, page_table: torch.Tensor, seq_lens: torch.Tensor, reserve: int, lengths: np.ndarray):
        self.page_table, self.seq_lens = page_table, seq_lens
        self.reserve = reserve             # new tokens every sequence has pages for
        self.lengths = lengths.copy()      # host copy of the prompt lengths
        self.steps = 0                     # tokens appended since the prefill
        self.shared_prefix = 0
        self.n_shared = 0
        self.group = None                  # rows with the same entry share their first n_shared pages (None: the whole batch)

    def get_seq_length(self) -> int:
        return int(self.lengths.max()) + self.steps


class CausalLMOutput(dict):
    """dict with attribute access (`out.logits`, `out['logits']`)."""
    __getattr__ = dict.get


class _Linear:
    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class _Core:
    """What `model.get_model()` returns: `.mm_projector`, `.embed_tokens`, `.config` and the adapter
    flags the reference's splice consults."""

    def __init__(self, outer):
        self._outer = outer
        self.config = outer.config
        self.clip_adapter = outer.config.clip_adapter
        self.hierarchy = outer.config.hierarchy

    @property
    def mm_projector(self):
        e = self._outer.engine
        return _Linear(e.proj_w, e.proj_b)

    @property
    def embed_tokens(self):
        return self._outer.engine.embed_tokens

    def get_input_embeddings(self):
        return self.embed_tokens


class RevisionLlamaForCausalLM:
    def __init__(self, config: RevisionConfig, state_dict: Dict[str, torch.Tensor],
                 clip_encoder_state: Optional[Dict[str, torch.Tensor]] = None):
        self.config = config
        self._sd = state_dict
        self._clip_sd = clip_encoder_state
        self.engine: Optional[Engine] = None
        self.device = torch.device("cpu")
        self.dtype = torch.bfloat16
        self.clip_encoder = None
        self.record_phase_events = False   # bench.py: CUDA events at the splice / prefill / decode boundaries of generate()
        self.last_phase_events = None
        self.share_prefix_compute = False  # also COMPUTE that prefix once (see _share_prefix_rows); opt-in
        self.share_prefix_pages = True     # map the KV pages of a prompt prefix common to the whole batch once (see _alloc_kv)
        self.decode_graphs = True          # replay chunks of greedy decode steps as CUDA graphs (engine.decode_chunk)

    # ---- placement (eval_nlq_negative.py:144-148 does `model.bfloat16().cuda()`)
    def bfloat16(self):
        return self

    def eval(self):
        return self

    def float(self):
        raise RvlError("fp32 execution is the CPU oracle's job (oracle/); this model only runs bf16 on sm_100")

    def to(self, *args, **kw):
        for a in list(args) + list(kw.values()):
            if a == torch.float32:
                return self.float()
            if isinstance(a, (str, torch.device)) and torch.device(a).type == "cuda":
                return self.cuda(torch.device(a).index)
        return self

    def cuda(self, device: Optional[int] = None):
        if self.engine is None:
            self.engine = Engine(self.config.engine_config(), device)
            self.engine.bind_state_dict(self._sd)
            self.device = self.engine.device
            self._sd = None        # the engine keeps the device copies alive
            if self._clip_sd is not None:
                from .clip_encoder import ClipEncoder
                self.clip_encoder = ClipEncoder(self.engine, self._clip_sd)
                self._clip_sd = None
        return self

    def get_model(self):
        return _Core(self)

    def get_input_embeddings(self):
        return self.engine.embed_tokens

    def _need_engine(self):
        if self.engine is None:
            raise RvlError("call .cuda() first: the model only runs on an sm_100 GPU (no CPU fallback)")

    # ---- visual adapter + splice ----------------------------------------------------------------
    def _visual_blocks(self, images, query_feats):
        """Returns (rows [n, 768 or hidden] bf16, n_visual per block, projected?)."""
        dev = self.device
        if isinstance(images, (list, tuple)):           # vtimellm_arch.py:102-109: list of [F_i, 768]
            n_vis = [int(im.shape[0]) for im in images]
            rows = torch.cat([im.reshape(-1, im.shape[-1]) for im in images], dim=0)
            return rows.to(dev, torch.bfloat16).contiguous(), n_vis, False
        if isinstance(images, WindowBank):               # hierarchy input with every distinct (window, query) pair stored once
            if self.clip_encoder is None:
                raise RvlError("a WindowBank needs the stage-2 ClipEncoder adapter (pass clip_encoder_state)")
            q_tok, q_mask = query_feats
            if isinstance(images.rows, (list, tuple)):          # prompts with different numbers of visual positions
                n_vis = [int(r.shape[0]) for r in images.rows]
                flat = torch.cat([r.reshape(-1) for r in images.rows])
            else:
                b, v = images.rows.shape
                n_vis, flat = [v] * b, images.rows.reshape(-1)
            cls_u = self.clip_encoder(images.windows.to(dev, torch.bfloat16).contiguous(), q_tok.to(dev, torch.bfloat16), q_mask.to(dev),
                                      images.text_index.to(dev, torch.int32).contiguous())              # [U, hidden] bf16
            return cls_u.index_select(0, flat.to(dev, torch.int64)).contiguous(), n_vis, True
        if images.dim() == 4:                            # hierarchy: [b, v, t, d] -> one CLS row per segment (:114-121)
            if self.clip_encoder is None:
                raise RvlError("4-D `images` need the stage-2 ClipEncoder adapter (pass clip_encoder_state)")
            b, v, t, d = images.shape
            q_tok, q_mask = query_feats
            feats = images.to(dev, torch.bfloat16).reshape(b * v, t, d).contiguous()
            # the reference repeats the query tokens/mask once per segment (vtimellm_arch.py:116-119); here every
            # segment of batch row i just points at text i
            seg_text = torch.arange(b, dtype=torch.int32, device=dev).repeat_interleave(v).contiguous()
            cls_rows = self.clip_encoder(feats, q_tok.to(dev, torch.bfloat16), q_mask.to(dev), seg_text)   # [b*v, hidden] bf16
            return cls_rows, [v] * b, True
        B, F, D = images.shape                           # stage 1: [B, F, 768] through the Linear projector (:125)
        return images.to(dev, torch.bfloat16).reshape(B * F, D).contiguous(), [F] * B, False

    def _share_prefix_rows(self, plan) -> None:
        """Opt-in (`share_prefix_compute`): the whole pages of a prompt prefix common to several sequences are computed ONCE.
        Sequences are grouped (`plan["group"]`: one group for the whole batch when only the text in front of <video> is
        common; one group per distinct visual input when `image_index` says which rows show the same segment - then the
        visual positions are common too).  The packed stream becomes [context of group 0 (P rows) | context of group 1 | ... |
        seq 0 from position P | seq 1 from position P | ...]; each context is one more sequence whose K/V go to the pages its
        group shares, and every real sequence names its group's context as its attention context (`seq_pos0` / `seq_ctx_row`
        of rvl_prefill).  Row-wise kernels (GEMMs, norms, SwiGLU) never see the difference and the attention tiles stay
        aligned to absolute positions, so every logit keeps its bits - (B - G) * P rows of work disappear (32 of 184 positions
        for the 1-hour sweep of one query; 128 of 184 for every further query on the same movie).  Off by default: the
        headline benchmark computes every segment in full, like the reference."""
        ps = self.engine.cfg.kv_page_size
        lengths = plan["lengths"].astype(np.int64)
        B = lengths.shape[0]
        group = np.asarray(plan.get("group", np.zeros(B, dtype=np.int64)), dtype=np.int64)
        uniq, first, inv = np.unique(group, return_index=True, return_inverse=True)      # leader = first member of each group
        G = uniq.shape[0]
        P = (min(int(plan["shared_prefix"]), int(lengths.min()) - 1) // ps) * ps
        if B < 2 or G == B or P < ps:
            return
        cu = plan["cu_seqlens"].astype(np.int64)
        own_start = np.concatenate([[G * P], G * P + np.cumsum(lengths - P)])          # first own row of sequence b; [-1] = total
        leader = np.zeros(B, dtype=bool)
        leader[first] = True
        def remap(dst):
            b = np.searchsorted(cu, dst, side="right") - 1
            pos = dst - cu[b]
            keep = (pos >= P) | leader[b]
            return keep, np.where(pos < P, inv[b] * P + pos, own_start[b] + pos - P)
        kt, nt = remap(plan["text_dst"].astype(np.int64))
        kv_, nv = remap(plan["vis_dst"].astype(np.int64))
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        plan["text_ids"], plan["text_dst"] = i32(plan["text_ids"][kt]), i32(nt[kt])
        plan["vis_src"], plan["vis_dst"] = i32(plan["vis_src"][kv_]), i32(nv[kv_])
        plan["cu_seqlens"] = i32(np.concatenate([np.arange(G) * P, own_start]))        # G + B sequences: the contexts first
        plan["ctx_len"] = P
        plan["ctx_groups"] = G
        plan["group_of"] = inv                                                          # dense group number of every sequence
        plan["group_leader"] = first

    def _splice(self, input_ids, attention_mask, images, query_feats, visual_memory=None, prefix_memory=None,
                share_compute: bool = False, image_index=None):
        """Projector + splice into a packed fp32 residual stream.  Returns (hidden [T, H], plan).
        `image_index` (int [B]): row b shows `images[image_index[b]]` - several prompts (queries) over the same segments
        without repeating the features; rows that share a segment form a group for `share_prefix_compute`."""
        eng = self.engine
        rows, n_vis, projected = self._visual_blocks(images, query_feats)
        idx_np = None
        if image_index is not None:
            if projected or isinstance(images, (list, tuple)) or visual_memory is not None:
                raise RvlError("image_index is defined for stage-1 `images` [S, F, D] (Linear projector, no visual_memory)")
            idx_np = np.asarray(image_index.detach().cpu().numpy() if isinstance(image_index, torch.Tensor) else image_index, dtype=np.int64)
            if idx_np.shape[0] != input_ids.shape[0] or idx_np.min() < 0 or idx_np.max() >= len(n_vis):
                raise RvlError("image_index needs one entry per prompt, each naming a row of `images`")
            n_src = n_vis
            n_vis = [n_src[i] for i in idx_np]                                          # visual rows of every prompt
        ids_np = input_ids.detach().cpu().numpy().astype(np.int64)
        am_np = None if attention_mask is None else attention_mask.detach().cpu().numpy().astype(bool)
        if visual_memory is not None:
            # <memory> streaming branch (vtimellm_arch.py:208-232): the -300 placeholder expands to
            # [embed_tokens(prefix_memory[b]) ; mm_projector(visual_memory[b])].  Here the placeholder is rewritten as the
            # prefix ids followed by one more visual placeholder, and the memory vectors join the projector GEMM as a second
            # visual block of the row - one GEMM for frames and memory, no concatenations.
            if projected or isinstance(images, (list, tuple)):
                raise RvlError("visual_memory is only defined for the stage-1 Linear projector ([B, F, D] `images`)")
            if attention_mask is not None:
                raise RvlError("visual_memory with an attention_mask is not part of the reference path")
            vm = visual_memory[:, None] if visual_memory.dim() == 2 else visual_memory       # [B, M, D]
            B, F = len(n_vis), n_vis[0]
            M = vm.shape[1]
            pm = prefix_memory.detach().cpu().numpy().astype(np.int64)                         # [B, P]
            new_ids = []
            for b in range(B):
                where = np.nonzero(ids_np[b] == constants.MEMORY_TOKEN_INDEX)[0]
                if where.shape[0] != 1 or (ids_np[b] == constants.IMAGE_TOKEN_INDEX).sum() != 1 or \
                        where[0] < np.argmax(ids_np[b] == constants.IMAGE_TOKEN_INDEX):
                    raise RvlError("visual_memory needs exactly one <video> followed by one <memory> placeholder per row")
                new_ids.append(np.concatenate([ids_np[b, :where[0]], pm[b], [constants.IMAGE_TOKEN_INDEX], ids_np[b, where[0] + 1:]]))
            ids_np = np.stack(new_ids)
            D = rows.shape[1]
            rows = torch.cat([rows.view(B, F, D), vm.to(self.device, torch.bfloat16)], dim=1).reshape(B * (F + M), D).contiguous()
            n_vis = [n for _ in range(B) for n in (F, M)]
        plan = plan_splice(ids_np, n_vis, am_np, self.config.tokenizer_model_max_length, constants.IMAGE_TOKEN_INDEX)
        pre = self._common_text_prefix(ids_np, am_np)
        plan["shared_prefix"] = min(pre, int(plan["lengths"].min()))
        plan["ctx_len"] = 0
        if idx_np is not None:
            # plan["vis_src"] counts the visual rows prompt by prompt: point them at the rows of the segment each prompt shows
            F = n_src[0]
            src = plan["vis_src"].astype(np.int64)
            plan["vis_src"] = np.ascontiguousarray(idx_np[src // F] * F + src % F, dtype=np.int32)
            # the same text up to a <video> placeholder at the same place in every row: prompts on the same segment also share
            # its visual positions
            if np.unique(idx_np).shape[0] < idx_np.shape[0] and pre < ids_np.shape[1] and \
                    bool((ids_np[:, pre] == constants.IMAGE_TOKEN_INDEX).all()):
                plan["group"] = idx_np
                plan["shared_prefix"] = min(pre + F, int(plan["lengths"].min()))
        if share_compute:
            self._share_prefix_rows(plan)
        dev = self.device
        T = int(plan["cu_seqlens"][-1])
        hidden = torch.empty((T, self.config.hidden_size), dtype=torch.float32, device=dev)
        h2d = lambda a: torch.from_numpy(a).to(dev, non_blocking=True)
        text_ids, text_dst = h2d(plan["text_ids"]), h2d(plan["text_dst"])
        vis_dst = h2d(plan["vis_dst"])
        if plan["vis_src"].shape[0] != rows.shape[0] or not np.array_equal(plan["vis_src"], np.arange(rows.shape[0])):
            rows = rows.index_select(0, h2d(plan["vis_src"]).long()).contiguous()   # truncation dropped some rows
        if projected:
            eng.splice_rows(rows, vis_dst, text_ids, text_dst, hidden)
        else:
            eng.project_splice(rows, vis_dst, text_ids, text_dst, hidden)
        return hidden, plan

    def _alloc_kv(self, lengths: np.ndarray, extra: int, shared_prefix: int = 0, group: Optional[np.ndarray] = None) -> KVState:
        """Page table of one live batch.  `shared_prefix` = number of leading prompt positions whose tokens are identical
        in every sequence (the system prompt in front of <video>): causal attention makes their K/V identical too, so
        the whole pages they fill are mapped to the SAME physical pages for all sequences.  Every sequence still
        computes and writes them at prefill (identical bits); what changes is that the 180 x 32 decode-attention CTAs
        of a step read one copy that stays in the 126 MB L2 instead of 180 copies from HBM.
        `group` (int [B]): the positions are identical only among rows of the same group (prompts on the same segment,
        `image_index`): each group gets its own copy of the shared pages."""
        eng = self.engine
        ps = eng.cfg.kv_page_size
        n_shared = (shared_prefix // ps) if (self.share_prefix_pages and len(lengths) > 1) else 0
        pages_per = [int(math.ceil((int(l) + extra) / ps)) for l in lengths]
        n_shared = min([n_shared] + [int(l) // ps for l in lengths])
        gid = np.zeros(len(lengths), dtype=np.int64) if group is None else np.unique(np.asarray(group), return_inverse=True)[1]
        n_groups = int(gid.max()) + 1 if len(lengths) else 0
        max_pages = max(pages_per)
        total = n_groups * n_shared + sum(n - n_shared for n in pages_per)
        eng.ensure_kv(total)
        table = np.zeros((len(lengths), max_pages), dtype=np.int32)
        nxt = n_groups * n_shared
        for i, n in enumerate(pages_per):
            table[i, :n_shared] = gid[i] * n_shared + np.arange(n_shared, dtype=np.int32)
            table[i, n_shared:n] = np.arange(nxt, nxt + n - n_shared, dtype=np.int32)
            nxt += n - n_shared
        dev = self.device
        kv = KVState(torch.from_numpy(table).to(dev), torch.from_numpy(lengths.astype(np.int32)).to(dev), extra, lengths)
        kv.shared_prefix = shared_prefix
        kv.n_shared = n_shared              # leading pages every row of a group maps to the same physical pages
        kv.group = None if group is None else np.asarray(group)
        return kv

    @staticmethod
    def _common_text_prefix(ids: np.ndarray, mask: Optional[np.ndarray]) -> int:
        """Leading positions that hold the same text token in every row (stops at the first placeholder / masked id)."""
        if ids.shape[0] < 2:
            return 0
        special = ids < 0
        if mask is not None:
            special = special | ~mask
        upto = int(np.where(special.any(axis=1), special.argmax(axis=1), ids.shape[1]).min())
        same = (ids[:, :upto] == ids[:1, :upto]).all(axis=0)
        return int(upto if same.all() else same.argmin())

    # ---- forward (vtimellm_llama.py:38-90) ---------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, images=None,
                visual_memory=None, prefix_memory=None, query_feats=None, return_dict=None, start_end_frame=None,
                iteration_step=None, logits_to_keep: int = 0, reserve_new_tokens: int = 64):
        self._need_engine()
        if (visual_memory is None) != (prefix_memory is None):
            raise RvlError("visual_memory and prefix_memory come together (vtimellm_arch.py:226-228)")
        if labels is not None or output_attentions:
            raise NotImplementedError("training-time outputs are out of scope (inference path only)")
        eng, cfg = self.engine, self.config
        dev = self.device
        with torch.cuda.device(dev):
            if past_key_values is not None and input_ids is not None and input_ids.shape[1] == 1:
                # decode step: position = tokens so far (vtimellm_arch.py:88-100)
                kv: KVState = past_key_values
                if kv.steps + 1 > kv.reserve:
                    raise RvlError("KV pages exhausted: pass a larger reserve_new_tokens to the prefill forward()")
                B = input_ids.shape[0]
                logits = torch.empty((B, cfg.vocab_size), dtype=torch.float32, device=dev)
                eng.decode_step(input_ids.reshape(B).to(dev, torch.int32).contiguous(), kv.seq_lens, kv.page_table, logits,
                                max_kv_len=kv.get_seq_length() + 1)
                kv.steps += 1
                return CausalLMOutput(logits=logits[:, None, :], past_key_values=kv)
            if inputs_embeds is not None:
                B, L, H = inputs_embeds.shape
                am = torch.ones((B, L), dtype=torch.bool) if attention_mask is None else attention_mask.bool().cpu()
                lengths = am.sum(1).numpy().astype(np.int32)
                hidden = inputs_embeds.to(dev)[am.to(dev)].float().contiguous()
                cu = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
            else:
                if images is None:
                    raise RvlError("forward() without `images` or `inputs_embeds` is not part of the scoring path")
                hidden, plan = self._splice(input_ids, attention_mask, images, query_feats, visual_memory, prefix_memory)
                lengths, cu = plan["lengths"], plan["cu_seqlens"]
            B = len(lengths)
            kv = self._alloc_kv(lengths, reserve_new_tokens, plan["shared_prefix"] if inputs_embeds is None else 0)
            cu_d = torch.from_numpy(cu).to(dev)
            Lmax, T = int(lengths.max()), int(cu[-1])
            if logits_to_keep == 1:
                last = torch.empty((B, cfg.vocab_size), dtype=torch.float32, device=dev)
                eng.prefill(hidden, cu_d, B, Lmax, kv.page_table, last, all_logits=False)
                return CausalLMOutput(logits=last[:, None, :], past_key_values=kv)
            flat = torch.empty((T, cfg.vocab_size), dtype=torch.float32, device=dev)
            eng.prefill(hidden, cu_d, B, Lmax, kv.page_table, flat, all_logits=True)
            logits = torch.zeros((B, Lmax, cfg.vocab_size), dtype=torch.float32, device=dev)   # right padded
            for b in range(B):
                logits[b, : lengths[b]] = flat[cu[b]: cu[b + 1]]
            return CausalLMOutput(logits=logits, past_key_values=kv)

    __call__ = forward

    # ---- generate (inference.py:45-59) ---------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids, images=None, query_feats=None, do_sample=False, temperature=1.0, num_beams=1,
                 max_new_tokens=1024, use_cache=True, visual_memory=None, prefix_memory=None, output_scores=False,
                 return_dict_in_generate=False, output_hidden_states=False, attention_mask=None,
                 eos_token_id="config", pad_token_id=None, stopping_criteria=None, seed: int = 0,
                 retire_finished: bool = False, mask_entropy_after_eos: bool = False, image_index=None, **unused):
        """`seed`: key of the Philox stream used when do_sample=True.  `retire_finished`: drop rows that emitted EOS from the
        decode batch (their remaining tokens are pad, as in the reference; their per-step entropies stop at EOS instead of
        continuing over pad inputs as the reference's do).  `mask_entropy_after_eos`: keep every row in the batch but report
        NaN for the entropies of the steps after a row's EOS - what the reference's one-call-per-prompt schedule (stage 2,
        eval_nlq_retrieval_e2e2.py:353-359) measures; the default keeps them, like a batched reference call does
        (eval_nlq_negative.py:287-297 takes the statistics over all steps of the batch).

        The loop looks at the EOS flags every `engine.DECODE_CHUNK` steps instead of every step (the reference's
        `unfinished_sequences.max() == 0` test, vtimellm_llama.py:359-362, costs one host synchronisation per token); steps
        that ran after the last row finished are trimmed from the outputs, so the result is the reference's.  Greedy decoding
        without per-step scores replays each chunk of steps as one CUDA graph (`decode_graphs`).
        `image_index` (int [B]): prompt b shows `images[image_index[b]]` ([S, F, D] stage-1 features) - several queries over
        the same segments in one batch; with `share_prefix_compute` the system text and the visual positions of a segment are
        computed once for all its prompts."""
        self._need_engine()
        if num_beams != 1:
            raise NotImplementedError("beam search is not part of the reference path (num_beams=1, inference.py:49)")
        if (visual_memory is None) != (prefix_memory is None):
            raise RvlError("visual_memory and prefix_memory come together (vtimellm_arch.py:226-228)")
        if images is None:
            raise RvlError("generate() needs `images` (pre-extracted CLIP features)")
        sampling = bool(do_sample) and temperature is not None and float(temperature) > 0.0
        eng, cfg = self.engine, self.config
        dev = self.device
        eos = cfg.eos_token_id if eos_token_id == "config" else eos_token_id
        pad = pad_token_id if pad_token_id is not None else (cfg.pad_token_id if cfg.pad_token_id is not None else (eos if eos is not None else 0))
        eos_arg = -1 if eos is None else int(eos)
        with torch.cuda.device(dev):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.record_phase_events else None
            if ev:
                ev[0].record()
            hidden, plan = self._splice(input_ids, attention_mask, images, query_feats, visual_memory, prefix_memory,
                                        share_compute=self.share_prefix_compute, image_index=image_index)
            if ev:
                ev[1].record()
            lengths, cu = plan["lengths"], plan["cu_seqlens"]
            B = len(lengths)
            room = cfg.max_position_embeddings - int(lengths.max())
            if room <= 0:
                raise RvlError(f"prompt of {int(lengths.max())} tokens exceeds max_position_embeddings={cfg.max_position_embeddings}")
            max_new = min(int(max_new_tokens), room)
            # KV pages: everything up front when small, else grow in chunks of 64 tokens as decoding proceeds
            chunk = max_new if max_new <= 64 else 64
            kv = self._alloc_kv(lengths, chunk, plan["shared_prefix"], plan.get("group"))
            # chunks of decode steps between two looks at the EOS flags; greedy chunks without per-step outputs are captured as
            # CUDA graphs over fixed-address buffers (engine.decode_chunk)
            chunked = not sampling and not retire_finished and not output_scores and max_new <= 64
            bufs = eng.decode_buffers(B, kv.page_table.shape[1]) if chunked else None
            cu_d = torch.from_numpy(cu).to(dev)
            P = int(plan["ctx_len"])
            self.last_shared_prefix = P         # positions of the prompt prefix computed once for the batch (0: none), for tests / tools
            if P:
                # G + B sequences: the context of every group (its K/V fill the pages the group shares), then every prompt from
                # position P on, naming its group's context rows
                ps = eng.cfg.kv_page_size
                if kv.n_shared < P // ps:
                    raise RvlError("share_prefix_compute needs the prefix pages mapped once per group (share_prefix_pages=True): "
                                   f"{P // ps} prefix pages are computed once but only {kv.n_shared} are shared")
                G = int(plan["ctx_groups"])
                lead = torch.from_numpy(np.ascontiguousarray(plan["group_leader"], dtype=np.int64)).to(dev)
                table = torch.zeros((G + B, kv.page_table.shape[1]), dtype=torch.int32, device=dev)
                table[:G, : P // ps] = kv.page_table.index_select(0, lead)[:, : P // ps]
                table[G:] = kv.page_table
                pos0 = torch.full((G + B,), P, dtype=torch.int32, device=dev)
                pos0[:G] = 0
                ctx_row = torch.zeros(G + B, dtype=torch.int32, device=dev)
                ctx_row[G:] = torch.from_numpy(np.ascontiguousarray(plan["group_of"] * P, dtype=np.int32)).to(dev)
                all_last = torch.empty((G + B, cfg.vocab_size), dtype=torch.float32, device=dev)
                eng.prefill(hidden, cu_d, G + B, int(lengths.max()), table, all_last, all_logits=False, seq_pos0=pos0, seq_ctx_row=ctx_row)
                if chunked:
                    logits = bufs["logits"]
                    logits.copy_(all_last[G:])
                else:
                    logits = all_last[G:].contiguous()
            else:
                logits = bufs["logits"] if chunked else torch.empty((B, cfg.vocab_size), dtype=torch.float32, device=dev)
                eng.prefill(hidden, cu_d, B, int(lengths.max()), kv.page_table, logits, all_logits=False)
            if ev:
                ev[2].record()
            probe = getattr(self, "debug_clock_probe", None)          # tools/phase_times.py: uint64 [3, 2] on the device
            if probe is not None:
                eng.lib.rvl_debug_sm_clock(probe[0].data_ptr(), torch.cuda.current_stream().cuda_stream)
            del hidden
            tokens = torch.full((max_new, B), int(pad), dtype=torch.int32, device=dev)
            entropies = torch.full((max_new, B), float("nan"), dtype=torch.float32, device=dev)
            unfinished = torch.ones(B, dtype=torch.int32, device=dev) if eos is not None else None
            scores: List[torch.Tensor] = []
            steps_run = 0
            poll = eng.DECODE_CHUNK
            if chunked:
                bufs["seq_lens"].copy_(kv.seq_lens)
                bufs["page_table"].copy_(kv.page_table)
                kv.seq_lens, kv.page_table = bufs["seq_lens"], bufs["page_table"]
                if unfinished is not None:
                    bufs["unfinished"].fill_(1)
                    unfinished = bufs["unfinished"]
                kv_bound = int(lengths.max()) + max_new
                t = 0
                while t < max_new - 1:
                    k = min(poll, max_new - 1 - t)
                    eng.decode_chunk(bufs, k, eos_arg, int(pad), unfinished is not None, kv_bound, graph=self.decode_graphs)
                    tokens[t: t + k].copy_(bufs["ring_tok"][:k])
                    entropies[t: t + k].copy_(bufs["ring_ent"][:k])
                    kv.steps += k
                    t += k
                    steps_run = t
                    if probe is not None and t == k:
                        eng.lib.rvl_debug_sm_clock(probe[1].data_ptr(), torch.cuda.current_stream().cuda_stream)
                    # one look at the EOS flags between two chunks (none after the last: the trim below sees everything)
                    if unfinished is not None and t < max_new - 1 and int(unfinished.sum().item()) == 0:
                        break
                else:
                    # the last token needs no decode step after it
                    eng.sample_greedy(logits, tokens[t], entropies[t], unfinished, eos_arg, int(pad))
                    steps_run = t + 1
            else:
                active = None                          # original row ids of the live decode batch once rows have been retired
                tok_t = ent_t = None
                for t in range(max_new):
                    n_live = logits.shape[0]
                    if active is None:
                        tok_t, ent_t = tokens[t], entropies[t]
                    else:
                        tok_t = torch.empty(n_live, dtype=torch.int32, device=dev)
                        ent_t = torch.empty(n_live, dtype=torch.float32, device=dev)
                    if sampling:
                        eng.sample_multinomial(logits, tok_t, float(temperature), seed, t, ent_t, unfinished, eos_arg, pad)
                    else:
                        eng.sample_greedy(logits, tok_t, ent_t, unfinished, eos_arg, pad)
                    if active is not None:
                        tokens[t].index_copy_(0, active, tok_t)
                        entropies[t].index_copy_(0, active, ent_t)
                    if output_scores:
                        if active is None:
                            scores.append(logits)
                        else:
                            full = torch.zeros((B, cfg.vocab_size), dtype=torch.float32, device=dev)
                            full.index_copy_(0, active, logits)
                            scores.append(full)
                    steps_run = t + 1
                    if t == max_new - 1:
                        break
                    n_unf = n_live
                    if unfinished is not None and (retire_finished or (t + 1) % poll == 0):
                        # vtimellm_llama.py:359-362 tests this every step; here every `poll` steps (every step only when rows
                        # are retired, which needs the host to know them) - the surplus steps are trimmed below
                        n_unf = int(unfinished.sum().item())
                        if n_unf == 0:
                            break
                    if t + 1 > kv.reserve:
                        kv = self._grow_kv(kv, kv.lengths, t + 1 + chunk)       # kv.lengths: the rows still in the batch
                    if retire_finished and unfinished is not None and n_unf < n_live:
                        # per-sequence retirement from the paged batch: only rows still generating go through the next step
                        live = torch.nonzero(unfinished, as_tuple=False).flatten()
                        active = live if active is None else active.index_select(0, live)
                        kv_live = KVState(kv.page_table.index_select(0, live).contiguous(), kv.seq_lens.index_select(0, live).contiguous(),
                                          kv.reserve, kv.lengths[live.cpu().numpy()])
                        kv_live.steps, kv_live.shared_prefix, kv_live.n_shared = kv.steps, getattr(kv, "shared_prefix", 0), kv.n_shared
                        kv_live.group = None if getattr(kv, "group", None) is None else kv.group[live.cpu().numpy()]
                        kv = kv_live
                        tok_t = tok_t.index_select(0, live).contiguous()
                        unfinished = unfinished.index_select(0, live).contiguous()
                        logits = torch.empty((live.shape[0], cfg.vocab_size), dtype=torch.float32, device=dev)
                    elif output_scores:
                        logits = torch.empty((n_live, cfg.vocab_size), dtype=torch.float32, device=dev)
                    eng.decode_step(tok_t.contiguous(), kv.seq_lens, kv.page_table, logits, max_kv_len=kv.get_seq_length() + 1)
                    kv.steps += 1
            if probe is not None:
                eng.lib.rvl_debug_sm_clock(probe[2].data_ptr(), torch.cuda.current_stream().cuda_stream)
            if ev:
                ev[3].record()
                self.last_phase_events = ev      # splice start, prefill start, decode start, end (read after a synchronize)
            n_steps = steps_run
            if eos is not None and steps_run > 0:
                # the reference stops right after the step in which the last row emitted EOS: trim what ran past it
                hit = tokens[:steps_run] == int(eos)                                    # [steps, B]
                first = torch.where(hit.any(dim=0), hit.to(torch.int32).argmax(dim=0), torch.full((B,), steps_run, device=dev))
                if mask_entropy_after_eos:
                    step_ix = torch.arange(steps_run, device=dev)[:, None]
                    entropies[:steps_run] = torch.where(step_ix > first[None, :], torch.full_like(entropies[:steps_run], float("nan")),
                                                        entropies[:steps_run])
                last = int(first.max().item())
                if last < steps_run:                                                     # every row finished
                    n_steps = last + 1
                    scores = scores[:n_steps]
            new_tokens = tokens[:n_steps].t().contiguous()
            ids_dev = input_ids.to(dev)
            sequences = torch.cat([ids_dev, new_tokens.to(ids_dev.dtype)], dim=1)     # prompt ids (placeholder echoed) + new
            if chunked:
                # hand back bookkeeping that does not alias the engine's reusable decode buffers
                kv.seq_lens, kv.page_table = kv.seq_lens.clone(), kv.page_table.clone()
            out = CausalLMOutput(sequences=sequences, scores=tuple(scores) if output_scores else None,
                                 entropies=entropies[:n_steps].t().contiguous(), prompt_lengths=torch.from_numpy(lengths.copy()),
                                 past_key_values=kv)
        if return_dict_in_generate:
            return out
        return sequences

    def _grow_kv(self, kv: KVState, lengths: np.ndarray, extra: int) -> KVState:
        """Re-page with more room, copying the live cache (rare: only for > 64 new tokens)."""
        eng = self.engine
        old_table, old_kv, old_pages = kv.page_table, eng._kv, eng.n_pages
        c = eng.cfg
        eng._kv = None
        eng.n_pages = 0
        new = self._alloc_kv(lengths, extra, getattr(kv, "shared_prefix", 0), getattr(kv, "group", None))
        per_old = old_pages * c.n_heads * c.kv_page_size * c.head_dim
        per_new = eng.n_pages * c.n_heads * c.kv_page_size * c.head_dim
        src, dst = old_kv.view(torch.bfloat16), eng._kv.view(torch.bfloat16)
        page_elems = c.n_heads * c.kv_page_size * c.head_dim
        n_old = old_table.shape[1]
        o_idx = old_table.long().reshape(-1)
        n_idx = new.page_table[:, :n_old].long().reshape(-1)
        for half in range(2 * c.n_layers):
            s = src[half * per_old:(half + 1) * per_old].view(old_pages, page_elems)
            d = dst[half * per_new:(half + 1) * per_new].view(eng.n_pages, page_elems)
            d[n_idx] = s[o_idx]
        new.seq_lens = kv.seq_lens
        new.steps = kv.steps
        return new

    # ---- construction helpers ----------------------------------------------------------------------
    @classmethod
    def from_synthetic(cls, synth_cfg, seed: int = 0, device: str = "cuda", clip_encoder: bool = False, **cfg_kw):
        from . import synthetic as syn
        sd = syn.make_llama_weights(synth_cfg, seed=seed, device=device)
        clip_sd = syn.make_clip_encoder_weights(synth_cfg.hidden, seed=seed, device=device) if clip_encoder else None
        cfg = RevisionConfig.from_synth(synth_cfg, clip_adapter=clip_encoder, clip_adapter_text=clip_encoder,
                                        hierarchy=clip_encoder, **cfg_kw)
        return cls(cfg, sd, clip_sd)
