"""Weight-streaming GEMM sweep (decode orientation): GB/s of weight bytes vs tokens / features."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import _cabi, synthetic as syn
from revisionllm_b200.engine import Engine, EngineConfig

eng = Engine(EngineConfig.from_synth(syn.TINY))
eng.ensure_workspace(512, 256)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def bench(M, N, K, mode=_cabi.GEMM_OUT_BF16, split_k=1, reps=20):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(4)]   # rotate weights: no L2 reuse across reps
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if mode == _cabi.GEMM_OUT_BF16 else torch.float32)
    for i in range(3):
        eng.gemm(A, Ws[i % 4], out=out, out_mode=mode, flags=_cabi.GEMM_FLAG_SWAP, split_k=split_k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        eng.gemm(A, Ws[i % 4], out=out, out_mode=mode, flags=_cabi.GEMM_FLAG_SWAP, split_k=split_k)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"M={M:4d} N={N:6d} K={K:6d} split_k={split_k} mode={mode}: {us:7.1f} us  {N*K*2/us/1e3:7.0f} GB/s  tiles={((N+127)//128)*split_k}")
for N in (12288, 18944, 22016, 32000):
    for M in (16, 32, 64, 128, 180, 256):
        bench(M, N, 4096)
for M in (16, 180):
    for sk in (1, 2, 4, 8):
        bench(M, 4096, 4096, mode=_cabi.GEMM_ADD_F32, split_k=sk)
        bench(M, 4096, 11008, mode=_cabi.GEMM_ADD_F32, split_k=sk)
