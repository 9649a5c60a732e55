"""Quick A/B of the weight-streaming GEMM at the decode shapes (run once per env setting, e.g. RVL_PROD=0/1/2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import _cabi, synthetic as syn
from revisionllm_b200.engine import Engine, EngineConfig

eng = Engine(EngineConfig.from_synth(syn.TINY))
eng.ensure_workspace(512, 256)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("RVL_"))
def bench(M, N, K, mode=_cabi.GEMM_OUT_BF16, split_k=1, reps=20):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(4)]
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if mode == _cabi.GEMM_OUT_BF16 else torch.float32)
    ref = (A.float() @ Ws[0].float().t())
    eng.gemm(A, Ws[0], out=out, out_mode=mode, flags=_cabi.GEMM_FLAG_SWAP, split_k=split_k)
    err = float((out.float() - ref).abs().max() / ref.abs().max()) if mode == _cabi.GEMM_OUT_BF16 else -1.0
    for i in range(3):
        eng.gemm(A, Ws[i % 4], out=out, out_mode=mode, flags=_cabi.GEMM_FLAG_SWAP, split_k=split_k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        eng.gemm(A, Ws[i % 4], out=out, out_mode=mode, flags=_cabi.GEMM_FLAG_SWAP, split_k=split_k)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"[{tag}] M={M:4d} N={N:6d} K={K:6d} split_k={split_k}: {us:7.1f} us  {N*K*2/us/1e3:7.0f} GB/s  err={err:.1e}", flush=True)
Ms = [int(x) for x in os.environ.get("AB_M", "16,180").split(",")]
for N in (12288, 22016, 32000):
    for M in Ms:
        bench(M, N, 4096)
for M in Ms:
    bench(M, 4096, 4096, mode=_cabi.GEMM_ADD_F32, split_k=4)
    bench(M, 4096, 11008, mode=_cabi.GEMM_ADD_F32, split_k=4)

# gate/up GEMM of the decode step: fused SwiGLU epilogue vs GEMM + swiglu kernel
def bench_gu(M, I=11008, K=4096, reps=20):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Ws = [torch.randn(2 * I, K, device="cuda").to(torch.bfloat16) for _ in range(4)]
    act = torch.zeros(M, I, device="cuda", dtype=torch.bfloat16)
    gu = torch.zeros(M, 2 * I, device="cuda", dtype=torch.bfloat16)
    def fused(i):
        eng.gemm(A, Ws[i % 4], out=act, flags=_cabi.GEMM_FLAG_SWAP | _cabi.GEMM_FLAG_SWIGLU | _cabi.GEMM_FLAG_W_CONST, ldc=I)
    def unfused(i):
        eng.gemm(A, Ws[i % 4], out=gu, flags=_cabi.GEMM_FLAG_SWAP | _cabi.GEMM_FLAG_W_CONST)
        eng.swiglu(gu)
    for name, fn in (("fused", fused), ("gemm+swiglu kernel", unfused)):
        for i in range(3):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        print(f"[{tag}] gate/up M={M}: {name:20s} {e0.elapsed_time(e1) / reps * 1e3:7.1f} us", flush=True)
for M in Ms:
    bench_gu(M)
