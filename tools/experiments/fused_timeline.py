"""Phase timeline of the experimental fused decode kernel (RVL_FUSED_DECODE=1): globaltimer stamps of the LAST fused launch
of a decode step (last layer: o, norm, gate|up, down, final norm, lm_head), relative to the earliest stamp."""
import ctypes as C, os, sys
os.environ["RVL_FUSED_DECODE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from revisionllm_b200 import synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM
B = int(sys.argv[1]) if len(sys.argv) > 1 else 180
cfg = syn.VICUNA_7B
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
eng = model.engine
ids = syn.make_prompt_ids(cfg, seed=2)
feats = syn.make_features(B, 100, 768, seed=1).cuda()
out = model(ids[None].expand(B, -1), images=feats, logits_to_keep=1, reserve_new_tokens=64)
kv = out.past_key_values
logits = out.logits[:, 0].contiguous()
tok = torch.zeros(B, dtype=torch.int32, device="cuda")
for t in range(3):
    eng.decode_step(tok, kv.seq_lens, kv.page_table, logits, max_kv_len=int(kv.lengths.max()) + t + 1)
torch.cuda.synchronize()
eng.lib.rvl_debug_fused_timestamps(1, None, 0)
eng.decode_step(tok, kv.seq_lens, kv.page_table, logits, max_kv_len=int(kv.lengths.max()) + 4)
torch.cuda.synchronize()
buf = np.zeros(160 * 6 * 4, dtype=np.uint64)
eng.lib.rvl_debug_fused_timestamps(0, buf.ctypes.data, buf.size)
t = buf.reshape(160, 6, 4).astype(np.int64)[:148]
t0 = t[t > 0].min()
names = ["o_proj", "norm", "gate|up", "down", "norm", "lm_head"]
print(f"B={B}: per phase, ns since the first stamp: [min..max over CTAs]")
for p in range(6):
    row = []
    for s, nm in enumerate(["released", "acc ready", "epi done", "arrived"]):
        v = t[:, p, s]
        v = v[v > 0] - t0
        row.append(f"{nm} {v.min() / 1e3:6.1f}..{v.max() / 1e3:6.1f} us" if v.size else f"{nm}   -")
    print(f"  {names[p]:8s} " + " | ".join(row))
