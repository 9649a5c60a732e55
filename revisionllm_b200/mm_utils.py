"""`tokenizer_image_token` - placeholder-aware tokenisation
(/root/reference/revisionllm/mm_utils.py:22-75): split the prompt at `<video>` (and `<memory>`),
tokenise the chunks, keep one BOS, insert -200 (and -300)."""
from __future__ import annotations

from typing import List

import torch

from .constants import DEFAULT_IMAGE_TOKEN, DEFAULT_MEMORY_TOKEN, IMAGE_TOKEN_INDEX, MEMORY_TOKEN_INDEX


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors=None):
    chunks = prompt.split(DEFAULT_IMAGE_TOKEN)
    with_memory = len(chunks) > 1 and DEFAULT_MEMORY_TOKEN in chunks[1]
    if with_memory:
        pieces = [chunks[0]] + chunks[1].split(DEFAULT_MEMORY_TOKEN)
    else:
        pieces = chunks
    tok = [tokenizer(p).input_ids for p in pieces]
    has_bos = bool(tok) and bool(tok[0]) and tok[0][0] == tokenizer.bos_token_id
    skip = 1 if has_bos else 0
    ids: List[int] = [tok[0][0]] if has_bos else []
    n_video_joined = 2 if with_memory else len(tok)
    for i in range(n_video_joined):
        ids.extend(tok[i][skip:])
        if i < n_video_joined - 1:
            ids.append(image_token_index)
    if with_memory:
        ids.append(MEMORY_TOKEN_INDEX)
        ids.extend(tok[2])          # the reference keeps this chunk's BOS (mm_utils.py:56)
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")
