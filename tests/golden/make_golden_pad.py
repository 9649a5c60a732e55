"""tests/golden/pad_sequences.npz: outputs of the REFERENCE's own `pad_sequences_1d`
(/root/reference/revisionllm/model/adapter/tensor_utils.py:5-53, loaded by path) on seeded inputs - ragged torch tensors,
ragged numpy arrays, nested lists with a torch and with a numpy dtype, a fixed length, and the driver's own call
(eval_nlq_negative.py:286: one query's token features repeated per window).  Run where /root/reference exists:

    python tests/golden/make_golden_pad.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_tensor_utils", "/root/reference/revisionllm/model/adapter/tensor_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def cases():
    g = torch.Generator().manual_seed(5)
    rng = np.random.default_rng(6)
    t3 = [torch.randn(n, 3, 4, generator=g) for n in (2, 5, 1)]
    n2 = [rng.standard_normal((n, 6)).astype(np.float32) for n in (4, 1, 3, 7)]
    lists = [[1, 2, 3], [1, 2], [3, 80, 7, 9]]
    q = torch.randn(12, 8, generator=g)
    return {
        "torch_ragged": (t3, dict(dtype=torch.float32)),
        "numpy_ragged": (n2, dict(dtype=np.float32)),
        "lists_torch": (lists, dict(dtype=torch.long)),
        "lists_numpy": (lists, dict(dtype=np.float32)),
        "fixed_length": (t3, dict(dtype=torch.float32, fixed_length=9)),
        "driver_call": (q[None].repeat(5, 1, 1), dict(dtype=q.dtype, device=q.device, fixed_length=None)),
    }


def main():
    out = {}
    for name, (seqs, kw) in cases().items():
        padded, mask = ref.pad_sequences_1d(seqs, **kw)
        out[name + "/padded"] = padded.numpy() if isinstance(padded, torch.Tensor) else padded
        out[name + "/mask"] = mask.numpy() if isinstance(mask, torch.Tensor) else mask
        out[name + "/is_torch"] = np.array(isinstance(padded, torch.Tensor))
    np.savez_compressed(os.path.join(HERE, "pad_sequences.npz"), **out)
    print(sorted(out))


if __name__ == "__main__":
    main()
