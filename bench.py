#!/usr/bin/env python
"""Benchmark of the stage-1 dense sweep (BASELINE.json metric: segments/sec, Vicuna-7B shape,
100-frame segments) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU fp32 path (oracle port)

A step = one pass of the hot path over one synthetic 1-hour MAD-shaped movie-query: 180 segments x 100
CLIP frames (768-d), 85 prompt ids (one <video> placeholder) -> L = 184, projector + splice + varlen
prefill + 16 greedy KV-cached decode steps + per-step entropy + CLIP cosine top-3 score per segment
(BASELINE.json configs[1]).  With N > 1 the SAME movie's 180 segments are dealt round-robin to the ranks and
one all-gather of the fixed-size per-segment records closes the step (BASELINE.json configs[2], strong
scaling: `value` = 180 segments / time of the slowest rank).  The weak-scaling figure (every rank sweeps
its own movie-query, the way the reference shards its eval by query, eval_nlq_negative.py:179-180) is
reported beside it as `weak_scaling`; `--scaling weak` makes it the headline instead.

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the
public sweep API with pinned host features and a device->host read of the records every step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "segments/sec (Vicuna-7B, 100-frame segs)"
N_SEG, N_FRAMES, NEW_TOKENS = 180, 100, 16
# algorithmic work per unit (BASELINE.md section 3 / SURVEY.md section 8d)
FLOP_PER_TOKEN = 12.952e9
WEIGHT_BYTES_PER_STEP = 13.214e9
KV_BYTES_PER_TOKEN = 0.524288e6


def workload_config(segments_per_rank_step: int, seq_len: int, world: int) -> dict:
    """`config` of the JSON line - the same for this repo's arm and for the reference arm."""
    return {"workload": "stage1_sweep_1h_movie (BASELINE.json configs[1]; configs[2] when sharded over N GPUs): 180 segments x 100 frames, "
                        "L=184, projector+splice+prefill+16 greedy decode steps+entropy+cosine top-3",
            "model": "Vicuna-7B shape (Llama-2-7B), random-init weights with four visually selected planted token chains",
            "segments_per_rank_step": segments_per_rank_step,
            "seq_len": seq_len, "new_tokens": NEW_TOKENS, "parallelism": f"segment-parallel dp{world}",
            "l2": "inputs larger than L2: 13.2 GB of weights are streamed every prefill/decode pass",
            "decoding": "greedy (north star); KV pages of the prompt prefix common to the batch are mapped once"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return None
        load = [c for c in sm if c >= 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_segment(weights_f32, cfg, feats_1, ids, new_tokens):
    """The reference's CPU fp32 path restated (oracle/): projector + splice + prefill + greedy decode of ONE segment."""
    from oracle import llama_ref, splice_ref
    shape = llama_ref.LlamaShape(cfg.hidden, cfg.n_layers, cfg.n_heads, cfg.head_dim, cfg.intermediate, cfg.vocab,
                                 cfg.rms_eps, cfg.rope_theta, cfg.adapter_dim)
    t0 = time.perf_counter()
    x = torch.stack(splice_ref.splice(weights_f32, ids[None], splice_ref.mm_projector_linear(weights_f32, feats_1)))
    toks, scores = llama_ref.greedy_decode(weights_f32, shape, x, new_tokens, stop_on_eos=False)
    return toks[0], torch.stack(scores)[:, 0], time.perf_counter() - t0


def host_weights_f32(sd_bf16):
    return {k: v.detach().to("cpu").float() for k, v in sd_bf16.items()}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  /root/reference does not exist on
    the GPU box and the decoder arithmetic lives in `transformers` (not vendored), so this is the oracle port
    (kind 'port'), fp32, all host threads, one segment per step."""
    if rank != 0:
        return
    from revisionllm_b200 import synthetic as syn
    cfg = syn.VICUNA_7B_VIS
    torch.set_num_threads(os.cpu_count() or 1)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sd = syn.make_llama_weights(cfg, seed=0, device=dev)
    w = host_weights_f32(sd)
    del sd
    if dev == "cuda":
        torch.cuda.empty_cache()
    feats = syn.make_features(N_SEG, N_FRAMES, cfg.adapter_dim, seed=1, class_cfg=cfg)
    ids = syn.make_prompt_ids(cfg, seed=2)
    times = []
    for i in range(args.warmup + args.steps):
        _, _, dt = oracle_segment(w, cfg, feats[i % N_SEG: i % N_SEG + 1].float(), ids, NEW_TOKENS)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = len(times) / total
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "segments/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(N_SEG, int(ids.shape[0]) - 1 + N_FRAMES, 1),
                           sample="CPU arm: each step scores 1 of the 180 segments (fp32 oracle port, all host threads)"),
            "cpu_baseline": {"value": val, "unit": "segments/s", "cores": cores, "kind": "port",
                             "sample": f"1 segment (L=184, {NEW_TOKENS} greedy tokens) per step, fp32, torch {torch.__version__}"},
            "e2e": {"value": val, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def synthetic_movie(cfg, n_frames_total: int, seed: int):
    """[T, 768] fp32 host features of a synthetic movie whose visual class changes every 500 frames."""
    from revisionllm_b200 import synthetic as syn
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_frames_total, cfg.adapter_dim, generator=g)
    dirs = syn.class_directions(cfg, 0)
    cls_of = syn.segment_classes((n_frames_total + 499) // 500, cfg.visual_classes)
    return x + cfg.class_amp * dirs[cls_of[torch.arange(n_frames_total) // 500]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): ONE movie's 180 segments over the N ranks; weak: one movie per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--segments", type=int, default=N_SEG)
    ap.add_argument("--no-stage2", action="store_true", help="skip the (untimed-in-value) stage-2 measurements")
    ap.add_argument("--ragged-videos", type=int, default=-1,
                    help="also time a VidChapters-shaped ragged batch (BASELINE.json configs[4]: stage 1 + stage 2) of this many videos, "
                         "sharded over the ranks; default: 1024 on 8 GPUs, none otherwise")
    ap.add_argument("--no-movie-e2e", action="store_true", help="skip the chained stage-1 -> stage-2 -> rank measurement")
    ap.add_argument("--no-multi-query", action="store_true", help="skip the 8-queries-per-movie measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from revisionllm_b200 import scoring, sweep, synthetic as syn
    from revisionllm_b200.model import RevisionConfig, RevisionLlamaForCausalLM
    cfg = syn.VICUNA_7B_VIS
    n_seg = args.segments
    if args.ragged_videos < 0:
        args.ragged_videos = 1024 if world == 8 else 0
    sd = syn.make_llama_weights(cfg, seed=0, device="cuda")
    keep_for_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline)
    sd_cpu_src = dict(sd) if keep_for_cpu else None
    model = RevisionLlamaForCausalLM(RevisionConfig.from_synth(cfg), sd).bfloat16().cuda(local_rank)
    del sd
    eng = model.engine
    dev = model.device
    strong = args.scaling == "strong"
    # the movie of the headline: the same on every rank (strong) or one per rank (weak)
    feats_host = syn.make_features(n_seg, N_FRAMES, cfg.adapter_dim, seed=1 + (0 if strong else rank), class_cfg=cfg).pin_memory()
    ids = syn.make_prompt_ids(cfg, seed=2)
    g = torch.Generator().manual_seed(3)
    cls_host = torch.randn(cfg.adapter_dim, generator=g).to(torch.bfloat16).pin_memory()
    feats_dev, cls_dev, ids_dev = feats_host.to(dev), cls_host.to(dev), ids.to(dev)
    seq_len = ids.shape[0] - 1 + N_FRAMES
    mine_np = sweep.shard_indices(n_seg, rank, world)
    mine_dev = torch.from_numpy(mine_np).to(dev)
    feats_mine_dev = feats_dev.index_select(0, mine_dev) if strong else feats_dev

    def resident_step(weak_feats=None):
        if weak_feats is None and strong:
            local = sweep.score_segments(model, feats_mine_dev, ids_dev, cls_dev, NEW_TOKENS, eos_token_id=None)
            return sweep.allgather_records(local, n_seg, rank, world)
        local = sweep.score_segments(model, feats_dev if weak_feats is None else weak_feats, ids_dev, cls_dev, NEW_TOKENS, eos_token_id=None)
        if world > 1:
            out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=dev)
            dist.all_gather_into_tensor(out, local)
            return out
        return local

    def e2e_step():
        if strong:
            res = sweep.stage1_sweep(model, feats_host, ids, cls_host, NEW_TOKENS, rank, world, eos_token_id=None)
            return res.records.cpu()
        local = sweep.score_segments(model, feats_host, ids, cls_host, NEW_TOKENS, eos_token_id=None)
        if world > 1:
            out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=dev)
            dist.all_gather_into_tensor(out, local)
            local = out
        return local.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """device time of `steps` calls, max over ranks, barrier + synchronize on both sides"""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)), out

    units_per_step = n_seg * (1 if strong else world)
    n_local = len(mine_np) if strong else n_seg
    model.record_phase_events = True          # four CUDA events per generate() call: splice / prefill / decode boundaries
    model.debug_clock_probe = torch.zeros((3, 2), dtype=torch.int64, device=dev)      # SM clock at decode start / after 8 steps / end
    for _ in range(max(args.warmup, 3)):
        rec = resident_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    dev_ms, rec = timed(resident_step, args.steps)
    launches = eng.launches - launches0
    pe = model.last_phase_events              # of the last timed step (already complete: the barrier synchronised)
    phase_ms = {"splice": pe[0].elapsed_time(pe[1]), "prefill": pe[1].elapsed_time(pe[2]), "decode": pe[2].elapsed_time(pe[3])}
    decode_mhz = [round(1e3 * c / max(n, 1)) for c, n in model.debug_clock_probe.cpu().tolist()]
    model.record_phase_events = False
    model.debug_clock_probe = None
    # ---- e2e: host features in, records out, every step
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rec_host = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    # ---- one extra profiled step: per-launch CUDA events on the launching stream (roofline numerators)
    model.decode_graphs = False               # per-launch events need per-launch enqueues: this one step runs without graph replay
    eng.profile(True)
    resident_step()
    torch.cuda.synchronize()
    eng.profile(False)
    model.decode_graphs = True
    pk = peaks()
    gemm = eng.profile_read(0)
    gemm_small = eng.profile_read(1)
    attn_p = eng.profile_read(2)
    attn_d = eng.profile_read(3)
    # decode step t reads K and V of (seq_len + t + 1) tokens per sequence, all layers
    attn_d_bytes = sum(n_local * KV_BYTES_PER_TOKEN * (seq_len + t + 1) for t in range(NEW_TOKENS - 1))
    # which measured peak: the GEMMs of a 1-hour sweep run back to back for hundreds of ms at the power-capped clock (the
    # sustained cuBLAS figure); a rank's share at N >= 4 is a < 100 ms burst between decode phases at nearly the full clock
    # (the burst figure).  Both are MEASURED_PEAKS.json numbers; the choice is by the time the GEMMs take in the step.
    sustained = gemm["ms"] >= 120.0
    tf_peak = pk["tf_sust"] if sustained else pk["tf_burst"]
    roofline = {
        "bound": "tensor", "kernel": "gemm_bf16_pair_kernel (token-major tcgen05 GEMMs of the prefill: qkv, o, gate|up + SwiGLU, down, projector)",
        "achieved": gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else None,
        "peak": tf_peak, "unit": "TFLOP/s",
        "peak_source": pk["src"] + (", sustained bf16 (GEMMs run >= 120 ms back to back in the step)" if sustained else
                                    ", burst bf16 (the GEMMs of this rank's share last < 120 ms)"),
        "frac": (gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 / tf_peak) if gemm["ms"] > 0 else None,
        "traffic": None, "launches": gemm["launches"], "ms_in_step": gemm["ms"],
        "decode_gemm": {"bound": "hbm", "achieved": gemm_small["bytes"] / (gemm_small["ms"] * 1e-3) / 1e9 if gemm_small["ms"] > 0 else None,
                        "peak": pk["hbm"], "unit": "GB/s", "ms_in_step": gemm_small["ms"], "launches": gemm_small["launches"],
                        "note": "per-launch event pairs in a step WITHOUT graph replay (an event between two kernels also removes their "
                                "programmatic-launch overlap): an upper bound on the in-graph time"},
        "decode_attention": {"bound": "hbm", "achieved": attn_d_bytes / (attn_d["ms"] * 1e-3) / 1e9 if attn_d["ms"] > 0 else None,
                             "peak": pk["hbm"], "unit": "GB/s", "ms_in_step": attn_d["ms"], "launches": attn_d["launches"],
                             "note": "algorithmic bytes = K/V of every cached token; the page of the prompt prefix shared by the batch (32 of "
                                     "~190 positions) is mapped once and served from L2, so DRAM bytes are ~17 % lower"},
        "prefill_attention_ms_in_step": attn_p["ms"],
    }
    # whole phases of the last timed step (CUDA events inside generate(), no per-launch events): what the user-visible
    # step is made of.  Decode: 15 steps, each streams the 13.2 GB of weights once and reads K/V of every cached token.
    dec_steps = NEW_TOKENS - 1
    dec_bytes = dec_steps * WEIGHT_BYTES_PER_STEP + attn_d_bytes
    # executed GEMM FLOPs: the o projection and the MLP of the LAST layer run on the last row of each sequence only
    # (csrc/engine.cu rvl_prefill; 2 * (H*H + 3*H*I) FLOP per skipped row), so they are not counted for the other rows
    skipped = n_local * (seq_len - 1) * 2.0 * (cfg.hidden * cfg.hidden + 3.0 * cfg.hidden * cfg.intermediate)
    tf_prefill = (n_local * seq_len * FLOP_PER_TOKEN - skipped) / (phase_ms["prefill"] * 1e-3) / 1e12
    roofline["phases"] = {
        "splice_ms": phase_ms["splice"], "prefill_ms": phase_ms["prefill"], "decode_ms": phase_ms["decode"],
        "prefill": {"bound": "tensor", "achieved": tf_prefill, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf_prefill / tf_peak,
                    "note": "all of prefill (GEMMs + attention + RMSNorm + RoPE/KV write + lm_head on last rows) against the executed GEMM FLOPs only"},
        "decode": {"bound": "hbm", "ms_per_step": phase_ms["decode"] / dec_steps, "achieved": dec_bytes / (phase_ms["decode"] * 1e-3) / 1e9,
                   "peak": pk["hbm"], "unit": "GB/s", "frac": dec_bytes / (phase_ms["decode"] * 1e-3) / 1e9 / pk["hbm"],
                   "sm_mhz_at_start_mid_end": decode_mhz,
                   "note": "15 decode steps (two CUDA-graph replays) + 16 sampling kernels inside the sweep; bytes = weights once per step + K/V "
                           "of every cached token.  The phase starts at the SM clock the power-capped prefill leaves behind and the clock "
                           "recovers over the ~100 ms it lasts (sm_mhz_at_start_mid_end, measured by a probe kernel)"},
    }
    # the same decode steps on their own (no prefill in front: the clock has recovered) - the steady-state figure of the kernels
    try:
        L0 = seq_len
        out0 = model(ids[None].expand(n_local, -1), images=feats_mine_dev, logits_to_keep=1, reserve_new_tokens=64)
        kv0 = out0.past_key_values
        bufs = eng.decode_buffers(n_local, kv0.page_table.shape[1])
        bufs["page_table"].copy_(kv0.page_table)
        bufs["logits"].copy_(out0.logits[:, 0])
        iso = []
        for rep in range(6):
            bufs["seq_lens"].copy_(kv0.seq_lens)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(2):
                eng.decode_chunk(bufs, eng.DECODE_CHUNK, -1, 0, False, L0 + 64)
            b.record()
            torch.cuda.synchronize()
            iso.append(a.elapsed_time(b) / (2 * eng.DECODE_CHUNK))
        iso_ms = sorted(iso[3:])[1]                                                   # reps 0-2: eager sights + capture
        iso_bytes = WEIGHT_BYTES_PER_STEP + n_local * KV_BYTES_PER_TOKEN * (L0 + eng.DECODE_CHUNK + 1)
        roofline["phases"]["decode_steady_state"] = {
            "bound": "hbm", "ms_per_step": iso_ms, "achieved": iso_bytes / (iso_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
            "frac": iso_bytes / (iso_ms * 1e-3) / 1e9 / pk["hbm"],
            "note": "16 decode steps of the same batch timed on their own (CUDA-graph replay, median of 3)"}
        del out0, kv0
    except Exception as e:
        roofline["phases"]["decode_steady_state"] = {"error": repr(e)[:200]}
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tr = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tr):
            t = json.load(open(tr))
            roofline["traffic"] = t.get("prefill_gemm_dram_bytes_per_launch")
            roofline["traffic_source"] = t.get("source")
            roofline["algorithmic_bytes_per_launch"] = t.get("prefill_gemm_algorithmic_bytes_per_launch")
            break
    # The power-capped SM clock differs from box to box (1.08 - 1.26 GHz seen), and MEASURED_PEAKS.json was taken on one of them:
    # the same library call it used (torch.matmul, bf16 8192^3, back to back until the power cap bites) is timed HERE as well,
    # for information - `frac` above stays relative to the driver-written peak.
    if rank == 0 and world == 1 and roofline["achieved"]:
        try:
            a8 = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
            b8 = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
            for _ in range(300):
                torch.matmul(a8, b8)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(1200):
                torch.matmul(a8, b8)
            c1.record()
            torch.cuda.synchronize()
            live = 1200 * 2.0 * 8192 ** 3 / (c0.elapsed_time(c1) * 1e-3) / 1e12
            roofline["cublas_sustained_on_this_box"] = {"tflops": live, "frac": roofline["achieved"] / live,
                                                        "note": "torch.matmul bf16 8192^3 x 1200 back to back after 300 warm-up calls, this process, after the timed region"}
            del a8, b8
        except Exception as e:
            roofline["cublas_sustained_on_this_box"] = {"error": repr(e)[:200]}
    if roofline["decode_gemm"]["achieved"]:
        roofline["decode_gemm"]["frac"] = roofline["decode_gemm"]["achieved"] / pk["hbm"]
    if roofline["decode_attention"]["achieved"]:
        roofline["decode_attention"]["frac"] = roofline["decode_attention"]["achieved"] / pk["hbm"]

    value = units_per_step * args.steps / (dev_ms * 1e-3)
    e2e_val = units_per_step * args.steps / e2e_s
    h2d = n_local * N_FRAMES * cfg.adapter_dim * 2 + ids.numel() * 8 + cfg.adapter_dim * 2
    d2h = int(rec_host.numel() * 4)
    dec_share = phase_ms["decode"] / (phase_ms["prefill"] + phase_ms["decode"] + phase_ms["splice"])
    line = {
        "metric": METRIC, "value": value, "unit": "segments/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(n_local, seq_len, world),
        "prefill_tokens_per_s": units_per_step * seq_len * args.steps / (dev_ms * 1e-3),
        "roofline": roofline,
        "e2e": {"value": e2e_val, "unit": "segments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "limiter": ("prefill GEMMs (gemm_bf16_pair_kernel) at the power-capped SM clock" if dec_share < 0.4 else
                    "decode weight streaming: every rank streams all 13.2 GB of weights per decode step for its share of the segments "
                    "(gemm_stream_pair_kernel / gemm_bf16_tcgen05_kernel + attn_decode_mma_kernel), "
                    f"{100 * dec_share:.0f} % of the step at {n_local} segments per rank"),
    }
    # ---- weak scaling beside the strong headline (N > 1): every rank sweeps its own movie-query
    if world > 1 and strong:
        try:
            feats_weak = syn.make_features(n_seg, N_FRAMES, cfg.adapter_dim, seed=1 + rank, class_cfg=cfg).to(dev)
            for _ in range(3):                 # new batch shape: eager sights, then the decode chunks are captured
                resident_step(feats_weak)
            wk_ms, _ = timed(lambda: resident_step(feats_weak), args.steps)
            line["weak_scaling"] = {"value": n_seg * world * args.steps / (wk_ms * 1e-3), "unit": "segments/s", "ms_per_step": wk_ms / args.steps,
                                    "segments_per_rank_step": n_seg, "note": "one movie-query per rank (the reference's split-by-query jobs)"}
            del feats_weak
        except Exception as e:
            line["weak_scaling"] = {"error": repr(e)[:200]}
    # ---- opt-in design point, reported beside the headline and NOT part of `value`: the page-aligned text prefix that all
    # 180 segments of a movie-query share (system prompt + "USER:", 32 of 184 positions) is projected and cached once instead
    # of 180 times (model.share_prefix_compute; bit-identical logits, tests/test_gpu_model.py).  The headline keeps computing
    # every segment in full, like the reference.
    extras = world == 1                       # the side measurements below belong to the 1-GPU line; N > 1 runs stay short
    try:
        if not extras:
            raise RuntimeError("reported by the 1-GPU run only")
        model.share_prefix_compute = True
        for _ in range(2):
            resident_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            rec_sp = resident_step()
        s1.record()
        barrier()
        sp_ms = max_over_ranks(s0.elapsed_time(s1))
        same = bool(torch.equal(sweep.unpack_records(rec_sp)["tokens"], sweep.unpack_records(rec)["tokens"]))
        line["shared_prefix_compute"] = {"value": units_per_step * args.steps / (sp_ms * 1e-3), "unit": "segments/s", "ms_per_step": sp_ms / args.steps,
                                         "tokens_identical_to_headline_run": same,
                                         "note": "opt-in: common 32-position prompt prefix computed once per step instead of once per segment"}
    except Exception as e:
        line["shared_prefix_compute"] = {"skipped": str(e)} if not extras else {"error": repr(e)[:200]}
    finally:
        model.share_prefix_compute = False
    # ---- several queries of the SAME movie in one pass (reported beside the headline, NOT part of `value`): the reference asks every
    # query of a movie about every window, one sweep per query (eval_nlq_negative.py:183-337; MAD has hundreds of queries per
    # movie), and SURVEY.md section 8e names batching the queries' sweeps as the way to keep the decode batch large when a movie is
    # spread over many GPUs.  8 queries x 180 segments: rows segment-major, features stored once (`image_index`), the system text
    # and the visual positions of a segment (128 of 184 positions) computed once per segment (`share_prefix_compute`), segments
    # dealt to the ranks, one all-gather.  Same tokens as one sweep per query.
    if strong and n_seg == N_SEG and not args.no_multi_query:
        try:
            from revisionllm_b200 import constants
            Qm = 8
            n_pre = int((ids == constants.IMAGE_TOKEN_INDEX).nonzero()[0])
            ids_q = ids[None].repeat(Qm, 1)
            for q in range(1, Qm):              # same system text, another query text behind <video> (the last ids stay: same planted chain)
                ids_q[q, n_pre + 1:-4] = torch.randint(3, cfg.vocab, (ids.shape[0] - n_pre - 5,), generator=torch.Generator().manual_seed(500 + q))
            cls_q = torch.randn(Qm, cfg.adapter_dim, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16).to(dev)
            model.share_prefix_compute = True
            mq = lambda: sweep.stage1_sweep_queries(model, feats_dev, ids_q, cls_q, NEW_TOKENS, rank, world, batch_segments=90, eos_token_id=None)
            for _ in range(3):
                rec_q = mq()
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                rec_q = mq()
            barrier()
            mq_ms = 1e3 * max_over_ranks(time.perf_counter() - t0) / 2
            tok_q = sweep.unpack_records(rec_q)["tokens"].view(n_seg, Qm, -1)
            tok_1 = sweep.unpack_records(rec)["tokens"][:n_seg]
            line["multi_query"] = {"queries": Qm, "segments": n_seg, "ms": mq_ms, "ms_per_query": mq_ms / Qm, "value": n_seg * Qm / (mq_ms * 1e-3),
                                   "unit": "segment-queries/s", "vs_one_sweep_per_query": (dev_ms / args.steps * Qm) / mq_ms,
                                   "shared_positions": int(model.last_shared_prefix), "rows_per_rank": int(len(sweep.shard_indices(n_seg, rank, world)) * Qm),
                                   "query0_tokens_identical_to_headline_run": bool((tok_q[:, 0] == tok_1).all()),
                                   "note": "8 queries on the headline's movie in one pass against 8 x the headline's step time; opt-in batching + shared visual context"}
        except Exception as e:
            line["multi_query"] = {"error": repr(e)[:300]}
        finally:
            model.share_prefix_compute = False
    # ---- the reference's own windowing of a 1-hour MAD movie (eval_nlq_negative.py:226-235: 250 frames per window, stride
    # half a window -> 57 windows, L = 334), reported beside the headline, not part of `value`
    try:
        if not extras:
            raise RuntimeError("reported by the 1-GPU run only")
        feats_mad = syn.make_features(57, 250, cfg.adapter_dim, seed=21, class_cfg=cfg).to(dev)
        for _ in range(3):                    # new batch shape: the third sight captures its decode chunks
            sweep.score_segments(model, feats_mad, ids_dev, cls_dev, NEW_TOKENS, eos_token_id=None)
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(args.steps):
            sweep.score_segments(model, feats_mad, ids_dev, cls_dev, NEW_TOKENS, eos_token_id=None)
        m1.record()
        torch.cuda.synchronize()
        mad_ms = m0.elapsed_time(m1) / args.steps
        line["mad_windows_57x250"] = {"ms_per_movie_query": mad_ms, "windows_per_s": 57 / (mad_ms * 1e-3), "prompt_len": seq_len - N_FRAMES + 250,
                                      "note": "one rank, 57 windows x 250 frames x 16 greedy tokens (the reference's MAD windowing)"}
        del feats_mad
    except Exception as e:
        line["mad_windows_57x250"] = {"skipped": str(e)} if not extras else {"error": repr(e)[:200]}
    # ---- stage 2 (BASELINE.json configs[3], reported beside the headline, not part of `value`): top-100 segments by
    # cosine score -> 250-frame windows through the ClipEncoder adapter (one CLS token per window) -> one ~180-token
    # prompt per zoom level (4, 2, 1), 16 greedy tokens each.  One query per rank.
    if not args.no_stage2 or not args.no_movie_e2e or args.ragged_videos > 0:
        from revisionllm_b200.clip_encoder import ClipEncoder
        model.clip_encoder = ClipEncoder(eng, syn.make_clip_encoder_weights(cfg.hidden, seed=0, device="cuda"))
    if not args.no_stage2 and extras:
        try:
            cos = sweep.unpack_records(rec)["cos"][:n_seg].contiguous()
            top = scoring.select_topk_segments(eng, cos, 100)
            wins = syn.make_features(100, 250, cfg.adapter_dim, seed=7).to(dev)           # the selected windows' 250 frames
            gq = torch.Generator().manual_seed(8)
            q_tok = torch.randn(1, 32, cfg.adapter_dim, generator=gq).to(torch.bfloat16)
            q_mask = torch.ones(1, 32)
            ids2 = syn.make_prompt_ids(cfg, seed=9)

            def stage2():
                return sweep.stage2_pass(model, wins, (q_tok, q_mask), ids2, grounding_windows=top.tolist(), batch=100,
                                         zooms=(4, 2, 1), max_new_tokens=NEW_TOKENS, perm_seed=0, eos_token_id=None)
            for _ in range(3):                # the third call captures the decode chunks as CUDA graphs
                stage2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r2 = stage2()
            torch.cuda.synchronize()
            dt2 = time.perf_counter() - t0
            line["stage2_top100"] = {"ms_per_query": 1e3 * dt2, "generate_calls": len(r2), "windows": 100, "frames_per_window": 250,
                                     "zooms": [4, 2, 1], "selected_first5": top[:5].tolist(),
                                     "note": "ClipEncoder (4 layers, d=768; each distinct window once) + splice of 100 CLS tokens + Vicuna-7B prefill/decode, 16 tokens per call"}
            # stage2_long_33: the top-33 windows of the same query (zooms 4 / 2 / 1 over chunks of 33 // zoom windows)
            top33 = scoring.select_topk_segments(eng, cos, 33)

            def stage2_33():
                return sweep.stage2_pass(model, wins[:33], (q_tok, q_mask), ids2, grounding_windows=top33.tolist(), batch=33,
                                         zooms=(4, 2, 1), max_new_tokens=NEW_TOKENS, perm_seed=0, eos_token_id=None)
            for _ in range(3):
                stage2_33()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r33 = stage2_33()
            torch.cuda.synchronize()
            line["stage2_top33"] = {"ms_per_query": 1e3 * (time.perf_counter() - t0), "generate_calls": len(r33), "windows": 33,
                                    "frames_per_window": 250, "zooms": [4, 2, 1]}
            # the same pass batched across 8 queries of one rank (sweep.stage2_pass_queries): the chunks of all queries share
            # the decode steps, so the 13 GB of weights stream once per step for 56 prompts instead of 7
            NQ = 8
            qs = [dict(windows=syn.make_features(100, 250, cfg.adapter_dim, seed=40 + k).to(dev),
                       query_feats=(torch.randn(1, 32, cfg.adapter_dim, generator=gq).to(torch.bfloat16), q_mask),
                       input_ids=syn.make_prompt_ids(cfg, seed=50 + k), grounding_windows=top.tolist(), perm_seed=k) for k in range(NQ)]

            def stage2_multi():
                return sweep.stage2_pass_queries(model, qs, batch=100, zooms=(4, 2, 1), max_new_tokens=NEW_TOKENS, eos_token_id=None)
            for _ in range(3):
                stage2_multi()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rq = stage2_multi()
            torch.cuda.synchronize()
            dtq = time.perf_counter() - t0
            line["stage2_top100"]["batched_across_queries"] = {"queries": NQ, "ms_per_query": 1e3 * dtq / NQ,
                                                               "generate_calls_batched": sum(len(r) for r in rq)}
            for it in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                sweep.stage2_pass(model, wins, (q_tok, q_mask), ids2, grounding_windows=top.tolist(), batch=100, zooms=(4, 2, 1),
                                  max_new_tokens=NEW_TOKENS, perm_seed=0, eos_token_id=None, dedup=False)
                torch.cuda.synchronize()
                line["stage2_top100"]["ms_per_query_stacked_repeats"] = 1e3 * (time.perf_counter() - t0)
        except Exception as e:      # stage 2 is reported, never allowed to take the headline number down
            line["stage2_top100"] = {"error": repr(e)[:200]}
    # ---- north_star "Target": ONE synthetic one-hour MAD-shaped movie through stage 1 (segment-sharded over the N ranks, one
    # all-gather) + answer parsing + stage-2 window selection + the stage-2 top-100 pass (calls dealt to the ranks) + merge / ranking, as one
    # chained call from HOST features (sweep.run_movie).  Windows: 200 feature frames sampled to 100, stride 100 -> 179 segments.
    if not args.no_movie_e2e:
        try:
            from functools import partial
            movie = synthetic_movie(cfg, 18000, seed=31).numpy()
            gq = torch.Generator().manual_seed(8)
            q_feats = (torch.randn(1, 32, cfg.adapter_dim, generator=gq).to(torch.bfloat16), torch.ones(1, 32))
            mc = sweep.MovieConfig(clip_length=200, num_frames=N_FRAMES, stage2_clip_length=200, stage2_num_frames=250, stride=5, batch=100,
                                   zooms=(4, 2, 1), max_new_tokens=NEW_TOKENS)
            ids_s2 = syn.make_prompt_ids(cfg, seed=9)

            def movie_call():
                return sweep.run_movie(model, movie, ids, cls_host, partial(syn.synthetic_answers, n_frames=N_FRAMES), (0.40, 0.45), mc,
                                       query_feats=q_feats, stage2_input_ids=ids_s2, detok_stage2=syn.synthetic_answers_stage2,
                                       rank=rank, world=world, eos_token_id=None)
            for _ in range(3):
                mres = movie_call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                mres = movie_call()
            barrier()
            mv_ms = 1e3 * max_over_ranks(time.perf_counter() - t0) / 2
            line["movie_e2e_ms"] = mv_ms
            phases = {}                            # one more run with the device synchronised at the phase boundaries (rank 0's view)
            sweep.run_movie(model, movie, ids, cls_host, partial(syn.synthetic_answers, n_frames=N_FRAMES), (0.40, 0.45), mc,
                            query_feats=q_feats, stage2_input_ids=ids_s2, detok_stage2=syn.synthetic_answers_stage2,
                            rank=rank, world=world, eos_token_id=None, timings=phases)
            barrier()
            line["movie_e2e"] = {"ms": mv_ms, "stage1_windows": int(mres.records.shape[0]), "frames_per_window": N_FRAMES,
                                 "stage1_answers_with_span": len(mres.clip_frames), "stage2_windows": len(mres.grounding_windows),
                                 "stage2_generate_calls": len(mres.stage2) if mres.stage2 else 0,
                                 "ranked_proposals": len(mres.ranked["windows"]) if mres.ranked else 0,
                                 "phases_ms": {k: round(v, 2) for k, v in phases.items()},
                                 "note": "wall clock, host features in -> ranked proposals out on rank 0: upload + window gather + stage 1 on N ranks + "
                                         "all-gather + parse + select + stage-2 top-100 (ClipEncoder + zooms 4/2/1; its independent generate() calls dealt to the N ranks, "
                                         "one all-gather of their records) + merge/rank kernel on rank 0; phases_ms: one extra run with the device synchronised "
                                         "at the phase boundaries"}
            # the same chain for 8 queries on that movie in one call (sweep.run_movie_queries: stage 1 in one pass with the visual
            # context of a window shared by its 8 prompts, stage 2 one GPU per query / batched across queries)
            if not args.no_multi_query:
                Qm = 8
                n_pre = int((ids == -200).nonzero()[0])
                ids_q = ids[None].repeat(Qm, 1)
                for q in range(1, Qm):
                    ids_q[q, n_pre + 1:-4] = torch.randint(3, cfg.vocab, (ids.shape[0] - n_pre - 5,), generator=torch.Generator().manual_seed(500 + q))
                cls_q = torch.randn(Qm, cfg.adapter_dim, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16)
                qf_q = [(torch.randn(1, 32, cfg.adapter_dim, generator=torch.Generator().manual_seed(700 + q)).to(torch.bfloat16), torch.ones(1, 32))
                        for q in range(Qm)]
                gts_q = [(0.05 + 0.1 * q, 0.10 + 0.1 * q) for q in range(Qm)]
                mcq = sweep.MovieConfig(clip_length=200, num_frames=N_FRAMES, stage2_clip_length=200, stage2_num_frames=250, stride=5, batch=100,
                                        zooms=(4, 2, 1), max_new_tokens=NEW_TOKENS, stage1_batch=90)
                model.share_prefix_compute = True
                try:
                    mq_call = lambda t=None: sweep.run_movie_queries(model, movie, ids_q, cls_q, partial(syn.synthetic_answers, n_frames=N_FRAMES), gts_q, mcq,
                                                                     query_feats=qf_q, stage2_input_ids=ids_s2, detok_stage2=syn.synthetic_answers_stage2,
                                                                     rank=rank, world=world, eos_token_id=None, timings=t)
                    for _ in range(3):
                        mq_res = mq_call()
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(2):
                        mq_res = mq_call()
                    barrier()
                    mq_e2e = 1e3 * max_over_ranks(time.perf_counter() - t0) / 2
                    ph_q = {}
                    mq_call(ph_q)
                    barrier()
                    line["movie_e2e_queries"] = {"queries": Qm, "ms": mq_e2e, "ms_per_query": mq_e2e / Qm, "vs_one_call_per_query": mv_ms * Qm / mq_e2e,
                                                 "stage2_queries_on_this_rank": sum(1 for r in mq_res if r.stage2 is not None),
                                                 "phases_ms": {k: round(v, 2) for k, v in ph_q.items()},
                                                 "note": "8 queries on the movie of movie_e2e in one sweep.run_movie_queries call (opt-in: shared visual context in "
                                                         "stage 1; stage 2 one GPU per query, batched across the queries of a rank), wall clock, max over ranks"}
                finally:
                    model.share_prefix_compute = False
        except Exception as e:
            line.setdefault("movie_e2e", {})["error"] = repr(e)[:300]
    # ---- BASELINE.json configs[4]: VidChapters-shaped ragged batch - videos of 1-60 min at 2 fps, 500 s windows with 250 s
    # stride sampled to <= 100 frames, queries of 8-32 tokens; stage 1 varlen-packed, windows dealt to the ranks by length
    # (sweep.shard_balanced), one all-gather of the records; then stage 2 (ClipEncoder + zooms 4/2/1 over each video's windows,
    # 16 windows per prompt), queries dealt to the ranks and batched across queries on each rank.
    if args.ragged_videos > 0:
        try:
            rng = np.random.default_rng(1)
            windows, qlens, video_of = [], [], []
            for v in range(args.ragged_videos):
                n_feat = int(rng.uniform(1, 60) * 60 * 2)
                if n_feat <= 1000:
                    spans = [(0, n_feat - 1)]
                else:
                    spans = [(i * 500, min(i * 500 + 1000, n_feat - 1)) for i in range(int(np.ceil(n_feat / 500)) - 1)]
                ql = int(rng.integers(8, 33))
                for (a0, b0) in spans:
                    windows.append(min(100, b0 - a0 + 1))
                    qlens.append(ql)
                    video_of.append(v)
            gw = torch.Generator().manual_seed(11)
            wins = [torch.randn(f, cfg.adapter_dim, generator=gw).to(torch.bfloat16).to(dev) for f in windows]
            Ltxt = 1 + 37 + 1 + 32 + 14
            rids = torch.zeros((len(wins), Ltxt), dtype=torch.int64)
            ram = torch.zeros((len(wins), Ltxt), dtype=torch.bool)
            for i, ql in enumerate(qlens):
                row = torch.cat([torch.tensor([1]), torch.randint(3, cfg.vocab, (37,), generator=gw), torch.tensor([-200]),
                                 torch.randint(3, cfg.vocab, (ql + 14,), generator=gw)])
                rids[i, : row.shape[0]] = row
                ram[i, : row.shape[0]] = True
            # stage-2 queries: one per video, its windows resampled to 100 frames each, a 16-token query text
            first = {}
            for i, v in enumerate(video_of):
                first.setdefault(v, []).append(i)
            s2_ids = syn.make_prompt_ids(cfg, seed=9)
            q_mask16 = torch.ones(1, 16)

            def s2_queries():
                qs = []
                for v, idxs in first.items():
                    if v % world != rank:
                        qs.append(dict(windows=torch.empty((len(idxs), 100, cfg.adapter_dim), dtype=torch.bfloat16), query_feats=None,
                                       input_ids=s2_ids, grounding_windows=list(range(len(idxs))), perm_seed=v))
                        continue
                    w = torch.stack([wins[i][torch.linspace(0, wins[i].shape[0] - 1, 100).long()] for i in idxs])
                    qf = (torch.randn(1, 16, cfg.adapter_dim, generator=torch.Generator().manual_seed(1000 + v)).to(torch.bfloat16), q_mask16)
                    qs.append(dict(windows=w, query_feats=qf, input_ids=s2_ids, grounding_windows=list(range(len(idxs))), perm_seed=v))
                return qs
            queries = s2_queries()

            def ragged():
                recs = sweep.ragged_sweep(model, wins, rids, ram, cls_dev, NEW_TOKENS, rank, world, max_tokens_per_batch=33120, eos_token_id=None)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                r2 = sweep.stage2_pass_queries(model, queries, batch=16, zooms=(4, 2, 1), max_new_tokens=NEW_TOKENS, eos_token_id=None,
                                               rank=rank, world=world)
                torch.cuda.synchronize()
                return recs, r2, t1
            ragged()
            barrier()
            t0 = time.perf_counter()
            recs, r2, t1 = ragged()
            barrier()
            t2 = time.perf_counter()
            tot_ms, s1_ms = 1e3 * max_over_ranks(t2 - t0), 1e3 * max_over_ranks(t1 - t0)
            line["vidchapters_ragged"] = {"videos": args.ragged_videos, "windows": len(wins),
                                          "prompt_tokens": int(sum(w + q + 52 for w, q in zip(windows, qlens))),
                                          "ms": tot_ms, "stage1_ms": s1_ms, "stage2_ms": tot_ms - s1_ms, "videos_per_s": args.ragged_videos / (tot_ms * 1e-3),
                                          "windows_per_s_stage1": len(wins) / (s1_ms * 1e-3), "new_tokens": NEW_TOKENS,
                                          "stage2_generate_calls_this_rank": sum(len(r) for r in r2 if r is not None),
                                          "note": "stage 1 over all windows of all videos (varlen-packed, weights replicated, windows balanced by length, one "
                                                  "all-gather) + stage 2 per video (queries dealt to the ranks, batched across queries); wall clock, max over ranks"}
        except Exception as e:
            line["vidchapters_ragged"] = {"error": repr(e)[:300]}
    # ---- CPU baseline beside it (rank 0, N=1): the oracle port on ONE segment, and full-size parity of that segment
    if keep_for_cpu:
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 0
        if avail > 48e9:
            torch.set_num_threads(os.cpu_count() or 1)
            w32 = host_weights_f32(sd_cpu_src)
            sd_cpu_src = None
            toks, sc, dt = oracle_segment(w32, cfg, feats_host[0:1].float(), ids, NEW_TOKENS)
            n_cpu, dt_cpu = 1, dt
            while dt_cpu < 10.0 and n_cpu < 8:                      # a bounded 10-30 s sample of the same workload
                dt_cpu += oracle_segment(w32, cfg, feats_host[n_cpu:n_cpu + 1].float(), ids, NEW_TOKENS)[2]
                n_cpu += 1
            out = model.generate(ids[None], images=feats_host[0:1], max_new_tokens=NEW_TOKENS, output_scores=True,
                                 return_dict_in_generate=True, eos_token_id=None)
            got = out["sequences"][0, ids.shape[0]:].cpu()
            gsc = torch.stack(out["scores"])[:, 0].cpu()
            rel = float((gsc.double() - sc.double()).abs().max() / sc.double().abs().max())
            line["cpu_baseline"] = {"value": n_cpu / dt_cpu, "unit": "segments/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{n_cpu} of {n_seg} segments one after the other (L={seq_len}, {NEW_TOKENS} greedy tokens each), "
                                              f"fp32 oracle, {dt_cpu:.1f} s"}
            line["parity_full_size"] = {"tokens_identical": bool(torch.equal(got.long(), toks.long())), "logit_max_rel_err": rel,
                                        "segment": 0, "segments_checked": 1, "oracle": "run live (the cpu_baseline sample)"}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "segments/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"skipped: only {avail / 1e9:.0f} GB host RAM available (fp32 7B needs 27 GB + headroom)"}
    # ---- full-size parity against the cached fp32 oracle (tests/golden/stage1_7b_vis.npz: 30 of the 180 segments, generated by
    # tests/golden/make_golden_7b.py): tokens of the timed run's own records + sampled logits of one more scored run
    fix = os.path.join(ROOT, "tests", "golden", "stage1_7b_vis.npz")
    if rank == 0 and world == 1 and strong and n_seg == N_SEG and os.path.exists(fix):
        try:
            gfx = np.load(fix)
            tok_run = sweep.unpack_records(rec.cpu())["tokens"][:, :NEW_TOKENS]
            segs = gfx["segments"].tolist()
            same = [tok_run[sg].tolist() == gfx["tokens"][n].tolist() for n, sg in enumerate(segs)]
            out = model.generate(ids[None].repeat(n_seg, 1), images=feats_dev, max_new_tokens=NEW_TOKENS, output_scores=True,
                                 return_dict_in_generate=True, eos_token_id=None)
            sc = torch.stack(out["scores"])
            worst = 0.0
            for n, sg in enumerate(segs):
                rows = sc[:, sg].float().cpu()
                got = torch.gather(rows, 1, torch.from_numpy(gfx["top_ids"][n].astype(np.int64)))
                err = (got - torch.from_numpy(gfx["top_vals"][n])).abs().amax(dim=1) / torch.from_numpy(gfx["row_absmax"][n])
                got_p = rows[:, torch.from_numpy(gfx["probe_ids"].astype(np.int64))]
                err_p = (got_p - torch.from_numpy(gfx["probe_vals"][n])).abs().amax(dim=1) / torch.from_numpy(gfx["row_absmax"][n])
                worst = max(worst, float(err.max()), float(err_p.max()))
            live = line.get("parity_full_size", {})
            line["parity_full_size"] = {"segments_checked": len(segs), "tokens_identical": int(sum(same)), "identical_fraction": sum(same) / len(segs),
                                        "logit_max_rel_err_per_row": worst, "tolerance": 3e-2, "distinct_token_chains": len({tuple(r) for r in gfx["tokens"].tolist()}),
                                        "oracle": "cached fp32 CPU oracle (tests/golden/stage1_7b_vis.npz); logits compared at each row's top-8 and 64 probe ids, "
                                                  "relative to the row's largest |logit|",
                                        "live_oracle_segment0": live}
            del sc, out
        except Exception as e:
            line.setdefault("parity_full_size", {})["fixture_error"] = repr(e)[:200]
    # ---- the same for stage 2 (BASELINE.json configs[3]; tests/golden/stage2_7b.npz from make_golden_7b_stage2.py): three prompts of
    # 100 ClipEncoder CLS tokens (zoom 1 / 2 / 4) over 100 windows x 250 frames, adapter weights from the CPU generator
    fix2 = os.path.join(ROOT, "tests", "golden", "stage2_7b.npz")
    if rank == 0 and world == 1 and not args.no_stage2 and os.path.exists(fix2):
        try:
            import importlib.util
            from revisionllm_b200.clip_encoder import ClipEncoder
            spec = importlib.util.spec_from_file_location("make_golden_7b_stage2", os.path.join(ROOT, "tests", "golden", "make_golden_7b_stage2.py"))
            gen2 = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(gen2)
            g2 = np.load(fix2)
            wins2 = syn.make_features(gen2.V, gen2.T, cfg.adapter_dim, seed=41)
            gq2 = torch.Generator().manual_seed(42)
            qt2 = torch.randn(1, gen2.LQ, cfg.adapter_dim, generator=gq2).to(torch.bfloat16)
            qm2 = torch.ones(1, gen2.LQ)
            qm2[0, gen2.LQ - 5:] = 0
            ids2 = syn.make_prompt_ids(cfg, seed=9)
            rows2 = [torch.arange(gen2.V), torch.arange(gen2.V // 2).repeat_interleave(2), torch.arange(gen2.V // 4).repeat_interleave(4)]
            keep_enc = model.clip_encoder
            model.clip_encoder = ClipEncoder(eng, syn.make_clip_encoder_weights(cfg.hidden, seed=0))
            n2, steps2 = g2["tokens"].shape
            o2 = model.generate(ids2[None].repeat(n2, 1), images=torch.stack([wins2[r] for r in rows2]), query_feats=(qt2.repeat(n2, 1, 1), qm2.repeat(n2, 1)),
                                max_new_tokens=steps2, output_scores=True, return_dict_in_generate=True, eos_token_id=None)
            model.clip_encoder = keep_enc
            tok2 = o2["sequences"][:, ids2.shape[0]:].cpu()
            sc2 = torch.stack(o2["scores"]).float().cpu()
            worst2 = 0.0
            for i in range(n2):
                got = torch.gather(sc2[:, i], 1, torch.from_numpy(g2["top_ids"][i].astype(np.int64)))
                err = (got - torch.from_numpy(g2["top_vals"][i])).abs().amax(dim=1) / torch.from_numpy(g2["row_absmax"][i])
                got_p = sc2[:, i][:, torch.from_numpy(g2["probe_ids"].astype(np.int64))]
                err_p = (got_p - torch.from_numpy(g2["probe_vals"][i])).abs().amax(dim=1) / torch.from_numpy(g2["row_absmax"][i])
                worst2 = max(worst2, float(err.max()), float(err_p.max()))
            line.setdefault("parity_full_size", {})["stage2"] = {
                "prompts_checked": int(n2), "tokens_identical": int(sum(tok2[i].tolist() == g2["tokens"][i].tolist() for i in range(n2))),
                "logit_max_rel_err_per_row": worst2, "tolerance": 3e-2,
                "oracle": "cached fp32 CPU oracle (tests/golden/stage2_7b.npz): ClipEncoder over 100 windows x 250 frames + 100 CLS tokens in a 7B prompt, zoom 1 / 2 / 4"}
        except Exception as e:
            line.setdefault("parity_full_size", {})["stage2"] = {"error": repr(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
