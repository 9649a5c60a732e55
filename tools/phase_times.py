"""Phase split of one stage-1 sweep step (180 segments, L=184, 16 tokens): splice+projector / prefill / decode / tail."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revisionllm_b200 import sweep, synthetic as syn
from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM
cfg = syn.VICUNA_7B
model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), syn.make_llama_weights(cfg, seed=0, device="cuda")).bfloat16().cuda()
model.record_phase_events = True
model.debug_clock_probe = torch.zeros((3, 2), dtype=torch.int64, device="cuda")     # SM clock at decode start / after 8 steps / at the end
feats = syn.make_features(180, 100, 768, seed=1).cuda()
ids = syn.make_prompt_ids(cfg, seed=2).cuda()
cls = torch.randn(768, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()
AB = os.environ.get("AB_ENV")          # e.g. AB_ENV=RVL_PDL (odd steps run with it set to 0) or AB_ENV=RVL_FULL_LAST_LAYER:1
AB_VAL = "0"
if AB and ":" in AB:
    AB, AB_VAL = AB.split(":")
acc = {0: [], 1: []}
graphs = {0: {}, 1: {}}                 # captured decode chunks bake the switches in: one graph cache per arm
for i in range(21 if AB else 6):
    if AB:
        if i % 2:
            os.environ[AB] = AB_VAL
        else:
            os.environ.pop(AB, None)
        model.engine.lib.rvl_reload_env()
        model.engine._dec_graphs = graphs[i % 2]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    sweep.score_segments(model, feats, ids, cls, 16, eos_token_id=None)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ev = model.last_phase_events
    print(f"step {i}: total {e0.elapsed_time(e1):7.2f} ms (host enqueue {1e3 * t_host:6.1f} ms) | before generate {e0.elapsed_time(ev[0]):6.2f} | "
          f"splice {ev[0].elapsed_time(ev[1]):6.2f} | prefill {ev[1].elapsed_time(ev[2]):7.2f} | decode {ev[2].elapsed_time(ev[3]):7.2f} | "
          f"tail {ev[3].elapsed_time(e1):6.2f} | SM MHz at decode start / mid / end "
          + " / ".join(f"{1e3 * c / max(n, 1):.0f}" for c, n in model.debug_clock_probe.cpu().tolist()), flush=True)
    if AB and i >= 9:                    # every arm has captured its graphs by then (third sight of the chunk shapes)
        acc[i % 2].append((ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), e0.elapsed_time(e1)))
if AB:
    for k, name in ((0, "default"), (1, AB + "=" + AB_VAL)):
        n = len(acc[k])
        print(f"{name:28s} prefill {sum(a[0] for a in acc[k]) / n:7.2f} ms  decode {sum(a[1] for a in acc[k]) / n:7.2f} ms  total {sum(a[2] for a in acc[k]) / n:7.2f} ms")
