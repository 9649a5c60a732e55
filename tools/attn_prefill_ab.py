"""Prefill attention: tcgen05 kernel (attention_tcgen05.cu) against the round-1 mma.sync kernel (attention.cu), interleaved in one
process, at the shapes of the sweep - 180 x L=184 (BASELINE configs[1]), 57 x L=334 (the reference's MAD windowing), one long
prompt - 32 heads x 128, plus a check of both against a float64 reference on a few (sequence, head) pairs."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from revisionllm_b200.engine import Engine, EngineConfig

eng = Engine(EngineConfig())
H, nh, d = 4096, 32, 128


def set_env(v):
    if v is None:
        os.environ.pop("RVL_ATTN_PREFILL", None)
    else:
        os.environ["RVL_ATTN_PREFILL"] = v
    eng.lib.rvl_reload_env()


for n_seq, L in ((180, 184), (57, 334), (8, 1484), (1, 4000)):
    lengths = [L - (i % 5) * 3 for i in range(n_seq)]                     # slightly ragged
    cu = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    T = int(cu[-1])
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = (torch.randn(T, 3 * H, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    cu_d = torch.from_numpy(cu).cuda()
    outs, times = {}, {}
    for rep in range(7):
        for name, env in (("tcgen05", None), ("mma.sync", "0")):
            set_env(env)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(4):
                o = eng.attn_prefill(qkv, cu_d, n_seq, max(lengths))
            b.record()
            torch.cuda.synchronize()
            if rep >= 2:
                times.setdefault(name, []).append(a.elapsed_time(b) / 4 * 1e3)
            outs[name] = o
    set_env(None)
    worst = {k: 0.0 for k in outs}
    for s in (0, n_seq // 2, n_seq - 1):
        for h in (0, 13, 31):
            x = qkv[cu[s]:cu[s + 1]].double()
            q, k, v = x[:, h * d:(h + 1) * d], x[:, H + h * d:H + (h + 1) * d], x[:, 2 * H + h * d:2 * H + (h + 1) * d]
            att = q @ k.t() / math.sqrt(d) + torch.triu(torch.full((len(x), len(x)), float("-inf"), device="cuda", dtype=torch.float64), 1)
            ref = torch.softmax(att, -1) @ v
            for name, o in outs.items():
                got = o[cu[s]:cu[s + 1], h * d:(h + 1) * d].double()
                worst[name] = max(worst[name], float((got - ref).abs().max() / ref.abs().max()))
    flops = sum(4.0 * l * l * d * nh / 2 for l in lengths)
    med = {k: sorted(v)[len(v) // 2] for k, v in times.items()}
    print(f"{n_seq:4d} x L={L:5d}: " + "  ".join(f"{k} {med[k]:8.1f} us ({flops / med[k] / 1e6:6.1f} TFLOP/s causal, rel err {worst[k]:.2e})" for k in med)
          + f"  speed-up {med['mma.sync'] / med['tcgen05']:.2f}x", flush=True)
