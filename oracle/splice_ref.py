"""Prompt assembly and the `<video>` embedding splice, restated (TEST INFRASTRUCTURE).

Follows:
  * /root/reference/revisionllm/conversation.py:51-60,253-263  (Vicuna v1 template, TWO style)
  * /root/reference/revisionllm/mm_utils.py:22-75              (tokenizer_image_token)
  * /root/reference/revisionllm/model/vtimellm_arch.py:81-299  (prepare_inputs_labels_for_multimodal)
  * /root/reference/revisionllm/model/vtimellm_arch.py:42      (mm_projector = Linear(adapter_input_dim, hidden))
  * /root/reference/revisionllm/model/adapter/tensor_utils.py:5-53 (pad_sequences_1d)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
MEMORY_TOKEN_INDEX = -300
DEFAULT_IMAGE_TOKEN = "<video>"
DEFAULT_MEMORY_TOKEN = "<memory>"

VICUNA_V1_SYSTEM = (
    "A chat between a curious user and an artificial intelligence assistant. "
    "The assistant gives helpful, detailed, and polite answers to the user's questions."
)


def vicuna_v1_prompt(query: str) -> str:
    """conv_templates['v1'] with one USER message and an open ASSISTANT turn
    (conversation.py:51-60 with sep=' ', sep2='</s>'; inference.py:31-34)."""
    return VICUNA_V1_SYSTEM + " " + "USER" + ": " + query + " " + "ASSISTANT" + ":"


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX) -> List[int]:
    """mm_utils.py:22-75: split on `<video>` (and `<memory>`), tokenize chunks,
    keep a single BOS, insert the placeholder ids."""
    image_chunks = prompt.split(DEFAULT_IMAGE_TOKEN)
    has_mem = len(image_chunks) > 1 and DEFAULT_MEMORY_TOKEN in image_chunks[1]
    if has_mem:
        prompt_chunks = [tokenizer(image_chunks[0]).input_ids]
        for mc in image_chunks[1].split(DEFAULT_MEMORY_TOKEN):
            prompt_chunks.append(tokenizer(mc).input_ids)
    else:
        prompt_chunks = [tokenizer(c).input_ids for c in image_chunks]

    def insert_separator(X, sep):
        return [e for sub in zip(X, [sep] * len(X)) for e in sub][:-1]

    ids: List[int] = []
    offset = 0
    if len(prompt_chunks) > 0 and len(prompt_chunks[0]) > 0 and prompt_chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        ids.append(prompt_chunks[0][0])
    if has_mem:
        for x in insert_separator(prompt_chunks[:2], [image_token_index] * (offset + 1)):
            ids.extend(x[offset:])
        ids.append(MEMORY_TOKEN_INDEX)
        ids.extend(prompt_chunks[2])
    else:
        for x in insert_separator(prompt_chunks, [image_token_index] * (offset + 1)):
            ids.extend(x[offset:])
    return ids


def pad_sequences_1d(seqs: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """tensor_utils.py:5-53: zero-pad along dim 0, fp32 validity mask."""
    lengths = [len(s) for s in seqs]
    out = torch.zeros((len(seqs), max(lengths)) + tuple(seqs[0].shape[1:]), dtype=seqs[0].dtype)
    mask = torch.zeros((len(seqs), max(lengths)), dtype=torch.float32)
    for i, s in enumerate(seqs):
        out[i, : lengths[i]] = s
        mask[i, : lengths[i]] = 1
    return out, mask


def mm_projector_linear(w: Dict[str, torch.Tensor], feats: torch.Tensor) -> torch.Tensor:
    """vtimellm_arch.py:42,125: nn.Linear(768 -> hidden) with bias."""
    return F.linear(feats.float(), w["model.mm_projector.weight"].float(), w["model.mm_projector.bias"].float())


def splice(
    w: Dict[str, torch.Tensor],
    input_ids: torch.Tensor,               # [B, Ltxt] int64 with one IMAGE_TOKEN_INDEX per row
    image_features: torch.Tensor,          # [B, F, hidden] already projected
    attention_mask: Optional[torch.Tensor] = None,
    max_length: Optional[int] = None,
    visual_memory: Optional[torch.Tensor] = None,    # [B, M, 768] (or [B, 768]) raw CLIP rows of the streaming memory
    prefix_memory: Optional[torch.Tensor] = None,    # [B, P] int64 text ids in front of the memory rows
) -> List[torch.Tensor]:
    """vtimellm_arch.py:149-244 (visual_memory=None branch): per row, drop
    padded ids, split at the placeholder(s), embed the text chunks, interleave
    the visual rows, truncate to `tokenizer_model_max_length`.  Returns the
    un-padded per-row embeddings (the reference then right-pads to the batch
    max, :246-276)."""
    emb = w["model.embed_tokens.weight"].float()
    out = []
    cur_image = 0
    if visual_memory is not None:
        # vtimellm_arch.py:208-232: [text ; video ; text ; embed(prefix_memory) ; mm_projector(visual_memory) ; text]
        vm = visual_memory[:, None] if visual_memory.dim() == 2 else visual_memory
        mem = torch.cat([emb[prefix_memory], mm_projector_linear(w, vm)], dim=1)          # [B, P + M, hidden]
        for b in range(input_ids.shape[0]):
            ids = input_ids[b]
            cut = [-1] + torch.where(ids == IMAGE_TOKEN_INDEX)[0].tolist() + torch.where(ids == MEMORY_TOKEN_INDEX)[0].tolist() + [ids.shape[0]]
            chunks = [emb[ids[cut[i] + 1: cut[i + 1]]] for i in range(len(cut) - 1)]
            parts = [chunks[0], image_features[b].float(), chunks[1], mem[b]] + ([chunks[2]] if len(chunks) == 3 else [])
            e = torch.cat(parts, dim=0)
            out.append(e[:max_length] if max_length is not None else e)
        return out
    for b in range(input_ids.shape[0]):
        ids = input_ids[b]
        if attention_mask is not None:
            ids = ids[attention_mask[b].bool()]
        pos = torch.where(ids == IMAGE_TOKEN_INDEX)[0].tolist()
        if len(pos) == 0:
            # :168-176 - text only; the image slot is consumed but contributes 0 rows
            e = emb[ids]
            out.append(e[:max_length] if max_length is not None else e)    # :239-243 truncates every row, text-only ones too
            cur_image += 1
            continue
        bounds = [-1] + pos + [ids.shape[0]]
        parts = []
        for i in range(len(bounds) - 1):
            parts.append(emb[ids[bounds[i] + 1: bounds[i + 1]]])
            if i < len(pos):
                f = image_features[cur_image]
                cur_image += 1
                if f.dim() == 1:
                    f = f[None]
                parts.append(f.float())
        e = torch.cat(parts, dim=0)
        if max_length is not None:
            e = e[:max_length]
        out.append(e)
    return out


def right_pad(embeds: List[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """vtimellm_arch.py:246-276 (padding_side='right'): stacked embeds, bool
    attention mask, position ids = arange(len) on the valid part."""
    B, Lm = len(embeds), max(e.shape[0] for e in embeds)
    H = embeds[0].shape[1]
    x = torch.zeros(B, Lm, H)
    m = torch.zeros(B, Lm, dtype=torch.bool)
    p = torch.zeros(B, Lm, dtype=torch.long)
    for i, e in enumerate(embeds):
        x[i, : e.shape[0]] = e
        m[i, : e.shape[0]] = True
        p[i, : e.shape[0]] = torch.arange(e.shape[0])
    return x, m, p


def decode_step_fixup(attention_mask: torch.Tensor, past_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """vtimellm_arch.py:88-100: at a 1-token step extend the mask to
    past_len+1 with ones and set position = sum(mask) - 1."""
    ext = torch.ones((attention_mask.shape[0], past_len + 1 - attention_mask.shape[1]), dtype=attention_mask.dtype)
    am = torch.cat((attention_mask, ext), dim=1)
    return am, am.sum(dim=1, keepdim=True) - 1
