"""`pad_sequences_1d` - the helper the stage-1 driver uses to batch the query token features
(/root/reference/revisionllm/eval/eval_nlq_negative.py:286: `pad_sequences_1d(query_feats[None].repeat(B, 1, 1), ...)`).

Same signature and results as /root/reference/revisionllm/model/adapter/tensor_utils.py:5-53: a list of sequences that
differ only in their first dimension becomes one zero-padded array plus a float32 validity mask; torch in -> torch out,
numpy (or nested lists with a numpy dtype) in -> numpy out.  Pinned by tests/golden/pad_sequences.npz, which the
reference's own function produced.  Note that `generate()` here does not need the repeated copies the driver builds
with it: `query_feats=(tokens [Q, Lq, 768], mask [Q, Lq])` is indexed per segment (model.WindowBank / seg_text_idx).
"""
from __future__ import annotations

import numpy as np
import torch


def _is_torch_dtype(dtype) -> bool:
    return isinstance(dtype, torch.dtype)


def pad_sequences_1d(sequences, dtype=torch.long, device=torch.device("cpu"), fixed_length=None):
    """-> (padded [n, L, ...], mask [n, L] float32 with 1 = valid).  L = `fixed_length` or the longest sequence."""
    use_torch = _is_torch_dtype(dtype)
    if isinstance(sequences[0], list):                                  # nested lists take their container from `dtype`
        sequences = [torch.tensor(s, dtype=dtype, device=device) if use_torch else np.asarray(s, dtype=dtype) for s in sequences]
    is_tensor = isinstance(sequences[0], torch.Tensor)
    if is_tensor != use_torch:
        raise AssertionError("dtype and input type does not match")     # the reference asserts the same condition
    lengths = [len(s) for s in sequences]
    L = max(lengths) if fixed_length is None else fixed_length
    shape = (len(sequences), L) + tuple(sequences[0].shape[1:])
    if is_tensor:
        padded = torch.zeros(shape, dtype=dtype, device=device)
        mask = torch.zeros(shape[:2], dtype=torch.float32, device=device)
    else:
        padded = np.zeros(shape, dtype=dtype)
        mask = np.zeros(shape[:2], dtype=np.float32)
    for row, (seq, n) in enumerate(zip(sequences, lengths)):
        padded[row, :n] = seq
        mask[row, :n] = 1
    return padded, mask
