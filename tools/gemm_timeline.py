"""Per-CTA role timestamps of one weight-streaming GEMM launch (debug hook rvl_debug_gemm_timestamps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from revisionllm_b200 import _cabi, synthetic as syn
from revisionllm_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig.from_synth(syn.TINY)); eng.ensure_workspace(512, 256)
lib = eng.lib
lib.rvl_debug_gemm_timestamps.argtypes = [C.c_int, C.c_void_p, C.c_int]
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): eng.gemm(A, W, out=out, flags=_cabi.GEMM_FLAG_SWAP)
torch.cuda.synchronize()
lib.rvl_debug_gemm_timestamps(1, None, 0)
eng.gemm(A, W, out=out, flags=_cabi.GEMM_FLAG_SWAP); torch.cuda.synchronize()
buf = np.zeros(160 * 8, dtype=np.uint64)
lib.rvl_debug_gemm_timestamps(0, buf.ctypes.data, buf.size)
t = buf.reshape(160, 8).astype(np.int64)
names = ["start", "prod_done", "mma_done", "acc_ready", "flags_seen|ld0", "epi_done", "published|ld1", "all_done"]
print(os.environ.get("RVL_EPI"), f"M={M} N={N} K={K}  (cycles relative to each CTA's start; 0 = not reached)")
for c in list(range(0, 3)) + [50, 100, 147]:
    if t[c, 0] == 0: continue
    print(f"cta {c:3d}: " + "  ".join(f"{n}={int(t[c, i] - t[c, 0]) if t[c, i] else 0:7d}" for i, n in enumerate(names) if i))
