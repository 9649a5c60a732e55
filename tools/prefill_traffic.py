"""The four prefill GEMMs of one layer (33120 tokens), one launch per rasterisation group size in argv, for an ncu DRAM-traffic pass:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16_pair \
      --csv --log-file gpurun_out/traffic.csv python tools/prefill_traffic.py 16 auto 64
Launch order in the csv: for each shape (qkv, o, gate|up, down): one warm-up launch, then one launch per group size."""
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
gms = sys.argv[1:] or ["auto"]
flush = torch.empty(1 << 28, dtype=torch.float32, device="cuda")       # 1 GiB: evict L2 between launches
for name, M, N, K, mode, flags in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    out = torch.zeros(M, N // 2 if flags else N, device="cuda", dtype=torch.bfloat16 if mode == _cabi.GEMM_OUT_BF16 else torch.float32)
    for gm in ["auto"] + gms:
        set_cfg(gm)
        flush.zero_()
        eng.gemm(A, W, out=out, out_mode=mode, flags=flags, ldc=out.shape[1])
        torch.cuda.synchronize()
    print(name, "done", flush=True)
