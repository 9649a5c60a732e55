"""Role timeline of CTA 0 of the tcgen05 prefill attention kernel (globaltimer stamps, csrc/attention_tcgen05.cu):
when the producer issued each K/V tile, when the MMA warp issued QK^T / PV, when softmax warp 2 saw the scores, finished its
exponentials, saw the previous PV done and published P."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from revisionllm_b200.engine import Engine, EngineConfig

n_seq, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (180, 184)
eng = Engine(EngineConfig())
H = 4096
cu = np.arange(n_seq + 1, dtype=np.int32) * L
qkv = (torch.randn(int(cu[-1]), 3 * H, device="cuda") * 1.5).to(torch.bfloat16)
cu_d = torch.from_numpy(cu).cuda()
for _ in range(3):
    eng.attn_prefill(qkv, cu_d, n_seq, L)
torch.cuda.synchronize()
eng.lib.rvl_debug_attn_timestamps(1, None, 0)
eng.attn_prefill(qkv, cu_d, n_seq, L)
torch.cuda.synchronize()
buf = np.zeros((4, 256, 4), dtype=np.uint64)
eng.lib.rvl_debug_attn_timestamps(0, buf.ctypes.data, buf.size)
t0 = int(buf[buf > 0].min())
rel = lambda v: (int(v) - t0) / 1e3 if v else float("nan")
print(f"{n_seq} x L={L}: CTA 0, times in us since its first stamp (slot 0 = earlier query tile, slot 1 = later)")
print(" tile | kv issued | QK0 issued  QK1 issued  PV0 issued  PV1 issued")
for t in range(18):
    print(f" {t:4d} | {rel(buf[0, t, 0]):9.2f} | " + " ".join(f"{rel(buf[1, t, k]):11.2f}" for k in range(4)))
print(" item | Q long issued  Q short issued | MMA warp starts the item")
for i in range(1, 8):
    print(f" {i:4d} | {rel(buf[0, 128 + i, 1]):13.2f} {rel(buf[0, 128 + i, 2]):15.2f} | {rel(buf[1, 128 + i, 0]):10.2f}")
for slot in (0, 1):
    print(f" slot {slot} tile | S ready  exps done  prevPV seen  P published")
    for t in range(14):
        print(f" {t:11d} | " + " ".join(f"{rel(buf[2 + slot, t, k]):9.2f}" for k in range(4)))
