"""GPU diagnostic battery for the tcgen05 GEMM: exact-integer inputs, structured so that a wrong smem
descriptor / K advance / TMEM lane mapping shows up as a recognisable error pattern.
Writes a human-readable log to stdout (run under gpurun, redirect to gpurun_out/)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from revisionllm_b200 import _cabi  # noqa: E402
from revisionllm_b200 import synthetic as syn  # noqa: E402
from revisionllm_b200.engine import Engine, EngineConfig  # noqa: E402


def main():
    print("device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    eng = Engine(EngineConfig.from_synth(syn.TINY))
    eng.ensure_workspace(512, 256)
    g = torch.Generator(device="cuda").manual_seed(0)

    def run(M, N, K, mode=_cabi.GEMM_OUT_F32, flags=0, tag="", a_mask=None):
        A = torch.randint(-3, 4, (M, K), device="cuda", generator=g).to(torch.bfloat16)
        W = torch.randint(-3, 4, (N, K), device="cuda", generator=g).to(torch.bfloat16)
        if a_mask is not None:
            A = A * a_mask.to(A.dtype)
        ref = A.float() @ W.float().t()
        t0 = time.time()
        out = eng.gemm(A, W, out_mode=mode, flags=flags)
        torch.cuda.synchronize()
        dt = time.time() - t0
        out = out.float()
        bad = (out != ref)
        nbad = int(bad.sum())
        print(f"[{tag}] M={M} N={N} K={K} flags={flags} mode={mode}: mismatches {nbad}/{out.numel()} maxabs {float((out-ref).abs().max()):.3f} ({dt*1e3:.1f} ms)")
        if nbad:
            rows = bad.any(1).nonzero().flatten().tolist()
            cols = bad.any(0).nonzero().flatten().tolist()
            print("   bad rows (first 24):", rows[:24], "... count", len(rows))
            print("   bad cols (first 24):", cols[:24], "... count", len(cols))
            print("   out[0,:8]", out[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())
            r = rows[0]
            print(f"   out[{r},:8]", out[r, :8].tolist(), f"ref[{r},:8]", ref[r, :8].tolist())
            # is it a permutation of rows / columns?
            if out.shape == ref.shape and M <= 256:
                match_rows = [(int((ref == out[i]).all(1).nonzero().flatten()[0]) if (ref == out[i]).all(1).any() else -1) for i in range(min(M, 16))]
                print("   out row i equals ref row:", match_rows)
        return nbad == 0

    ok = True
    ok &= run(128, 256, 64, tag="1tile-1kblock")
    # only the j-th 16-wide K slice non-zero: checks the +32 B descriptor advance
    for j in range(4):
        m = torch.zeros(1, 64, device="cuda")
        m[:, 16 * j:16 * j + 16] = 1
        ok &= run(128, 256, 64, tag=f"kslice{j}", a_mask=m)
    ok &= run(128, 32, 64, tag="BN32")
    ok &= run(128, 64, 64, tag="BN64")
    ok &= run(128, 128, 64, tag="BN128")
    ok &= run(128, 256, 512, tag="kloop8")
    ok &= run(128, 256, 4096, tag="kloop64 (ring wraps)")
    ok &= run(512, 1024, 256, tag="multi-tile")
    ok &= run(200, 264, 192, tag="tails")
    ok &= run(128, 256, 64, mode=_cabi.GEMM_OUT_BF16, tag="bf16 out")
    ok &= run(64, 512, 256, flags=_cabi.GEMM_FLAG_SWAP, tag="swap")
    ok &= run(180, 4096, 4096, flags=_cabi.GEMM_FLAG_SWAP, tag="swap decode-like")
    ok &= run(33120 // 4, 4096, 4096, mode=_cabi.GEMM_OUT_BF16, tag="prefill-like")
    print("ALL OK" if ok else "SOME FAILED")
    # quick perf probe
    for (M, N, K) in ((33120, 12288, 4096), (33120, 22016, 4096), (33120, 4096, 11008), (33120, 4096, 4096)):
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(2):
            eng.gemm(A, W, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.gemm(A, W, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        e0.record()
        for _ in range(5):
            torch.matmul(A, W.t(), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 5
        print(f"perf M={M} N={N} K={K}: ours {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TFLOP/s | torch.matmul {ms_t:.3f} ms = {2*M*N*K/ms_t/1e9:.0f} TFLOP/s")
        del A, W, out
    # decode-like weight streaming
    for (M, N, K) in ((180, 12288, 4096), (180, 22016, 4096), (180, 32000, 4096), (23, 22016, 4096)):
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(2):
            eng.gemm(A, W, out=out, flags=_cabi.GEMM_FLAG_SWAP)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.gemm(A, W, out=out, flags=_cabi.GEMM_FLAG_SWAP)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        e0.record()
        for _ in range(10):
            torch.matmul(A, W.t(), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
        print(f"decode M={M} N={N} K={K}: ours {ms*1e3:.1f} us = {N*K*2/ms/1e6:.0f} GB/s | torch {ms_t*1e3:.1f} us = {N*K*2/ms_t/1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
