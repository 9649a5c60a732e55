"""Prefill GEMM shapes (33120 tokens) timed for several rasterisation group sizes (RVL_GROUP_M is read per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import _cabi, synthetic as syn
from revisionllm_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig.from_synth(syn.TINY))


def set_cfg(cfg):
    """cfg: "auto", a bare group size ("16"), or VAR=VALUE pairs joined by commas ("RVL_GROUP_M=16")."""
    for k in ("RVL_GROUP_M",):
        os.environ.pop(k, None)
    if cfg != "auto":
        if "=" not in cfg:
            os.environ["RVL_GROUP_M"] = cfg
        else:
            for kv in cfg.split(","):
                k, v = kv.split("=")
                os.environ[k] = v
    _cabi.load().rvl_reload_env()


T, H, I = 33120, 4096, 11008
shapes = [("qkv", T, 3 * H, H, _cabi.GEMM_OUT_BF16, 0), ("o", T, H, H, _cabi.GEMM_ADD_F32, 0), ("gate|up+swiglu", T, 2 * I, H, _cabi.GEMM_OUT_BF16, _cabi.GEMM_FLAG_SWIGLU),
          ("down", T, H, I, _cabi.GEMM_ADD_F32, 0)]
for name, M, N, K, mode, flags in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    out = torch.zeros(M, N // 2 if flags else N, device="cuda", dtype=torch.bfloat16 if mode == _cabi.GEMM_OUT_BF16 else torch.float32)
    gms = sys.argv[1:] or ["16", "auto"]
    tot = {g: 0.0 for g in gms}
    rounds = 8
    for r in range(rounds + 1):                      # interleaved: the power-capped clock drifts within seconds
        for gm in (gms if r % 2 == 0 else gms[::-1]):
            set_cfg(gm)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                eng.gemm(A, W, out=out, out_mode=mode, flags=flags, ldc=out.shape[1])
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                tot[gm] += e0.elapsed_time(e1) / 5
    for gm in gms:
        ms = tot[gm] / rounds
        print(f"{name:16s} cfg={gm:34s} {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
