"""TEST INFRASTRUCTURE ONLY (CPU oracle) - never imported by the product path.

Restatement of the reference's sampling rule (revisionllm/model/vtimellm_llama.py:312-338, called with do_sample=True,
temperature=0.05 from revisionllm/inference.py:47-48): scores = logits / T, probs = softmax(scores), next ~ multinomial(probs).
torch.multinomial's random stream is an implementation detail (it differs between devices and torch versions), so the
CUDA path and this oracle share a documented one instead: Philox4x32-10 (Salmon et al. 2011), key = 64-bit seed,
counter = (row, step, 0, 0), u = (x0 >> 8) / 2^24, draw = first index whose inclusive CDF of exp((x - max) / T) exceeds
u * total.  Pinned by the Random123 known-answer vectors (tests/test_oracle_golden.py).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF

KNOWN_ANSWERS = [   # (counter, key, output) - Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((_MASK,) * 4, (_MASK,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(counter: Tuple[int, int, int, int], key: Tuple[int, int]) -> Tuple[int, int, int, int]:
    c = list(counter)
    k0, k1 = key
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c[3] ^ k1) & _MASK, p0 & _MASK]
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return tuple(c)


def uniform(seed: int, step: int, row: int) -> float:
    x0 = philox4x32_10((row, step, 0, 0), (seed & _MASK, (seed >> 32) & _MASK))[0]
    return (x0 >> 8) / 16777216.0


def multinomial_draw(logits: np.ndarray, temperature: float, seed: int, step: int) -> Tuple[List[int], List[float]]:
    """logits [B, V] -> (tokens, slack): float64 inverse CDF; slack[b] = distance of u * total to the nearest CDF edge
    relative to total (a float32 implementation may land on the neighbouring index when the slack is ~1e-6)."""
    toks, slack = [], []
    for b in range(logits.shape[0]):
        x = logits[b].astype(np.float64)
        w = np.exp((x - x.max()) / temperature)
        cdf = np.cumsum(w)
        target = uniform(seed, step, b) * cdf[-1]
        i = int(np.searchsorted(cdf, target, side="right"))
        i = min(i, logits.shape[1] - 1)
        toks.append(i)
        edges = [abs(cdf[i] - target)] + ([abs(target - cdf[i - 1])] if i > 0 else [])
        slack.append(min(edges) / cdf[-1])
    return toks, slack
