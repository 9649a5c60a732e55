"""Rows-per-tile experiment for the weight-streaming GEMM: run with RVL_BM=<rows> in the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import _cabi, synthetic as syn
from revisionllm_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig.from_synth(syn.TINY))
eng.ensure_workspace(512, 256)
def bench(M, N, K, reps=20):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(4)]
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    for i in range(3):
        eng.gemm(A, Ws[i % 4], out=out, flags=_cabi.GEMM_FLAG_SWAP)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        eng.gemm(A, Ws[i % 4], out=out, flags=_cabi.GEMM_FLAG_SWAP)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    bm = int(os.environ.get("RVL_BM", "0"))
    print(f"BM={bm:3d} M={M:4d} N={N:6d}: {us:7.1f} us {N*K*2/us/1e3:7.0f} GB/s tiles={(N + bm - 1)//bm if bm else -1}")
for N in (12288, 22016, 32000, 4096):
    for M in (16, 180):
        bench(M, N, 4096)
