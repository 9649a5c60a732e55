"""Would tensor-parallel decode beat the replicated design for ONE movie on 8 GPUs?

Strong scaling at N = 8 leaves 22 - 23 segments per rank; every rank still streams all 13.2 GB of weights per decode step
(3.2 ms per step in the steady state).  Tensor-parallel decode would keep all 180 sequences on every rank and give each rank
1/8 of every weight (column-parallel qkv / gate|up, row-parallel o / down with an all-reduce of the [180, 4096] partial sums
after each).  This script measures, on ONE GPU, the part of such a step that needs no second GPU: the 4 x 32 + 1 weight-streaming
GEMMs at their 1/8 shard shapes with 180 tokens, back to back inside a CUDA graph (so launch overhead is the replayed-graph
overhead, as in production), next to the same chain at full size with 23 tokens.  Add the measured attention / RMSNorm time of a
step and 64 all-reduces to the first number to get a lower bound for a tensor-parallel step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import _cabi
from revisionllm_b200.engine import Engine, EngineConfig

eng = Engine(EngineConfig())
eng.ensure_workspace(256, 256)
H, I, V, L = 4096, 11008, 32000, 32
FL = _cabi.GEMM_FLAG_SWAP | _cabi.GEMM_FLAG_W_CONST


def chain(tokens, tp):
    """(weights, activations, outputs) of one layer's four GEMMs + lm_head at 1 / tp of the weight"""
    g = torch.Generator(device="cuda").manual_seed(1)
    rnd = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    i_shard = ((I // tp + 15) // 16) * 16
    shapes = [("qkv", 3 * H // tp, H, _cabi.GEMM_OUT_BF16, 0), ("o", H, H // tp, _cabi.GEMM_OUT_F32, 0),
              ("gate|up", 2 * i_shard, H, _cabi.GEMM_OUT_BF16, _cabi.GEMM_FLAG_SWIGLU), ("down", H, i_shard, _cabi.GEMM_OUT_F32, 0)]
    ops = []
    for name, N, K, mode, fl in shapes:
        W = [rnd(N, K) for _ in range(4)]                                   # four copies rotate: 4 x the layer's bytes > L2 at full size
        A = rnd(tokens, K)
        out = torch.empty((tokens, N // 2 if fl else N), device="cuda", dtype=torch.bfloat16 if mode == _cabi.GEMM_OUT_BF16 else torch.float32)
        ops.append((name, W, A, out, mode, fl))
    head = (rnd(V // tp, H), rnd(tokens, H), torch.empty((tokens, V // tp), device="cuda", dtype=torch.float32))
    return ops, head


def step(ops, head):
    for l in range(L):
        for name, W, A, out, mode, fl in ops:
            eng.gemm(A, W[l % 4], out=out, out_mode=mode, flags=FL | fl, ldc=out.shape[1])
    eng.gemm(head[1], head[0], out=head[2], out_mode=_cabi.GEMM_OUT_F32, flags=FL)


for label, tokens, tp in (("replicated, 23 sequences per rank, full weights", 23, 1), ("tensor-parallel shard (1/8 of every weight), 180 sequences", 180, 8),
                          ("replicated, 180 sequences, full weights (1-GPU sweep)", 180, 1)):
    ops, head = chain(tokens, tp)
    step(ops, head)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        gr.capture_begin(capture_error_mode="thread_local")
        step(ops, head)
        gr.capture_end()
    torch.cuda.current_stream().wait_stream(s)
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(4):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 4)
    wbytes = sum(W[0].numel() * 2 for _, W, *_ in ops) * L + head[0].numel() * 2
    ms = sorted(ts)[len(ts) // 2]
    print(f"{label:62s}: {ms:6.3f} ms for the {4 * L + 1} GEMMs of a step ({ms * 1e3 / (4 * L + 1):5.1f} us each), {wbytes / 1e9:6.2f} GB of weights "
          f"-> {wbytes / ms / 1e6:6.0f} GB/s", flush=True)
