#!/usr/bin/env python
"""bench.py -- the BASELINE.json headline metric on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path

metric: training tokens/sec, Llama-3-8B QLoRA (NF4 double-quant, r=16 on the 7 projections),
bf16, seq 2048, batch 4 per GPU (BASELINE.json configs[1]; DDP over N GPUs = configs[3]).
A "step" = H2D of the batch (e2e leg) -> forward -> backward -> LoRA-grad all-reduce -> AdamW on
the LoRA bucket.  Synthetic token batches, random-init weights of the named architecture (no
network for datasets / checkpoints).

Two timed legs of K steps each after W warm-ups, CUDA events bracketed by barrier + synchronize,
max over ranks:
  value : inputs already resident in HBM
  e2e   : through the public API (model(input_ids=..., labels=...)) with the batch copied from
          pinned host memory every step and the loss read back to the host
Extra keys: roofline (dominant kernel = the tcgen05 multi-segment GEMM, timed live with CUDA
events on the launching stream), cpu_baseline (the reference's CPU path = torch eager fp32 HF
Llama + plain LoRA on a bounded sample, host cores stated), clocks, gpu_launches, peak VRAM.
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

METRIC = "train_tokens_per_sec"
UNIT = "tokens/s"
MODEL = "llama-3-8b"
SEQ, BS, RANK_R = 2048, 4, 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu-reference"],
                    help="ours | reference (the reference's CPU path, this tier's reference arm) | gpu-reference "
                         "(the reference's own GPU kernels on the B200: BASELINE.md B1/B3)")
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--layers", type=int, default=None, help="debug only: fewer layers (INVALID as a bench)")
    ap.add_argument("--seq", type=int, default=SEQ)
    ap.add_argument("--bs", type=int, default=BS)
    ap.add_argument("--preset", default=None, choices=["cfg2", "cfg3", "cfg5"],
                    help="BASELINE.json configs[1] (default) / [2] Mistral-7B r=32 seq4096 bs2 sliding window / "
                         "[4] Gemma-2-9B r=16 seq8192 bs1 -- cfg3/cfg5 are parity/coverage runs, not the headline")
    ap.add_argument("--rank", type=int, default=RANK_R)
    ap.add_argument("--sliding-window", type=int, default=None)
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--gradient-checkpointing", nargs="?", const="true", default=None, choices=["true", "unsloth"],
                    help="recompute every decoder layer in the backward; `unsloth` also parks the layer inputs in "
                         "pinned host memory (off: the BASELINE configs fit a B200)")
    ap.add_argument("--tiled-mlp", type=int, default=0, help="run the MLP over N token shards (recomputed shard-wise)")
    ap.add_argument("--no-keep-dequant", action="store_true",
                    help="re-dequantise the NF4 weights in the backward (the reference's schedule) instead of "
                         "keeping the forward's 16-bit expansion resident until the layer's backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-tokens", type=int, default=None,
                    help="tokens of the CPU arm's bounded sample (default: one row of --seq tokens)")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip timing the reference's GPU path (Triton kernels + cuBLAS LoRA schedule) beside ours")
    a = ap.parse_args()
    if a.preset == "cfg3":
        a.model, a.seq, a.bs, a.rank = "mistral-7b-v0.3", 4096, 2, 32
        a.sliding_window = a.sliding_window or 2048      # v0.3 ships sliding_window=null; exercise the band
        a.no_cpu_baseline = True
    elif a.preset == "cfg5":
        a.model, a.seq, a.bs, a.rank = "gemma-2-9b", 8192, 1, 16
        a.no_cpu_baseline = True
    return a


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load": samples in the upper half of the power range
        if sm:
            thr = (max(power) + min(power)) / 2
            load = [s for s, p in zip(sm, power) if p >= thr] or sm
            return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": max(power)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons), "samples": 0}


# ---------------------------------------------------------------------------------------------
# the reference's CPU path: torch eager fp32, stock HF Llama + plain LoRA
# ---------------------------------------------------------------------------------------------
def host_threads():
    """A FIXED thread count for the CPU arm: the cores this process may actually use (affinity
    mask and cgroup CPU quota), capped at 64 -- torch's CPU GEMMs stop scaling beyond that and an
    over-subscribed quota collapses them (round 1: 1.3 tokens/s with 128 threads on a box whose
    quota was lower).  No per-run probing: the same box always gets the same number."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, 64))


def cpu_reference_tokens_per_sec(model_name, seq_tokens, r=RANK_R, threads=None, reps=3, depths=(1, 2)):
    """Times forward+backward of the stock HuggingFace implementation on the host cores for a
    BOUNDED sample: `seq_tokens` tokens (one row at the workload's full sequence length) through
    models with `depths` decoder layers of the named architecture, `reps` timed repetitions each
    after one warm-up, MEDIAN taken; per-layer and embed/lm_head/CE costs separated by differencing
    and extrapolated to the full depth.  Imports nothing from unsloth_b200 (the arm must not load
    the product).  Returns (tokens/s of the full model, description, threads, spread)."""
    import torch
    from torch import nn
    from unsloth_b200.model_configs import CONFIGS, hf_config

    threads = threads or host_threads()
    torch.set_num_threads(threads)
    full_layers = CONFIGS[model_name]["num_hidden_layers"]

    class PlainLoRA(nn.Module):                      # y = x W^T + s (x A^T) B^T
        def __init__(self, lin, r, alpha):
            super().__init__()
            self.lin = lin
            self.A = nn.Parameter(torch.empty(r, lin.in_features).uniform_(-0.01, 0.01))
            self.B = nn.Parameter(torch.zeros(lin.out_features, r))
            self.s = alpha / r
            lin.weight.requires_grad_(False)

        def forward(self, x):
            return self.lin(x) + self.s * (x @ self.A.t()) @ self.B.t()

    def step_times(n_layers):
        from transformers import AutoModelForCausalLM
        cfg = hf_config(model_name, n_layers)
        cfg._attn_implementation = "sdpa"
        torch.manual_seed(0)
        m = AutoModelForCausalLM.from_config(cfg).float()
        for p in m.parameters():
            p.requires_grad_(False)
        for layer in m.model.layers:
            for parent in (layer.self_attn, layer.mlp):
                for nme in ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"):
                    if hasattr(parent, nme):
                        setattr(parent, nme, PlainLoRA(getattr(parent, nme), r, r))
        ids = torch.randint(0, cfg.vocab_size, (1, seq_tokens))
        ts = []
        for i in range(reps + 1):
            t0 = time.perf_counter()
            out = m(input_ids=ids, labels=ids)
            out.loss.backward()
            if i > 0:
                ts.append(time.perf_counter() - t0)
        del m
        return ts

    times = {d: step_times(d) for d in depths}
    med = {d: statistics.median(t) for d, t in times.items()}
    d0, d1 = depths[0], depths[-1]
    per_layer = max((med[d1] - med[d0]) / max(d1 - d0, 1), 1e-9)
    head = max(med[d0] - per_layer * d0, 0.0)
    full = head + per_layer * full_layers
    spread = {str(d): [round(min(t), 3), round(max(t), 3)] for d, t in times.items()}
    desc = ("stock HF %s (transformers, torch eager, fp32, device=cpu) + plain LoRA r=%d, fwd+bwd of 1x%d "
            "tokens; %s-layer models, %d reps each after 1 warm-up, medians %s s (min/max %s); per-layer cost "
            "extrapolated to %d layers + embed/lm_head/CE; %d threads (fixed rule: usable cores capped at 64)"
            % (model_name, r, seq_tokens, "/".join(map(str, depths)), reps,
               "/".join("%.2f" % med[d] for d in depths), json.dumps(spread), full_layers, threads))
    return seq_tokens / full, desc, threads, spread


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    v, desc, threads, spread = cpu_reference_tokens_per_sec(args.model, args.cpu_sample_tokens or args.seq,
                                                             r=args.rank, reps=max(3, min(args.steps, 5)))
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(args.bs * args.seq / v * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "%s QLoRA r=%d seq%d bs%d/GPU (reference CPU path: torch eager fp32)" %
                                   (args.model, args.rank, args.seq, args.bs)},
            "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": desc, "spread_s": spread},
            "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from unsloth_b200 import _lib as L
    from unsloth_b200.ddp import FlatLoRABucket, init_distributed
    from unsloth_b200.kernels import utils as KU
    from unsloth_b200.patch import build_qlora_model, lora_parameters

    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    extra = {}
    if args.sliding_window:
        extra["sliding_window"] = args.sliding_window
    if args.seq > 8192:
        extra["max_position_embeddings"] = args.seq
    from unsloth_b200.kernels import utils as KU
    if args.no_keep_dequant:
        KU.set_keep_dequant(False)
    model = build_qlora_model(args.model, r=args.rank, lora_alpha=args.rank, device=dev, seed=3407,
                              num_hidden_layers=args.layers,
                              gradient_checkpointing={"true": True, "unsloth": "unsloth", None: False}[args.gradient_checkpointing],
                              tiled_mlp=args.tiled_mlp, **extra)
    bucket = FlatLoRABucket(lora_parameters(model), lr=2e-4, weight_decay=0.01)
    bucket.broadcast_params(0)
    V = model.config.vocab_size
    total_steps = args.warmup + args.steps
    g = torch.Generator().manual_seed(1234 + rank)
    host_ids = torch.randint(0, V, (total_steps, args.bs, args.seq), generator=g).pin_memory()
    host_lab = host_ids.clone()
    host_lab[torch.rand(host_lab.shape, generator=g) < 0.1] = -100   # ~10% ignored
    host_lab = host_lab.pin_memory()
    dev_ids, dev_lab = host_ids.to(dev), host_lab.to(dev)

    def eager_step(ids, lab):
        bucket.begin_step()                 # zero_grad + batched LoRA cast refresh
        try:
            out = model(input_ids=ids, labels=lab)
            out.loss.backward()
        finally:
            bucket.end_backward()           # batched accumulation of the LoRA gradients
        bucket.all_reduce_grads()
        bucket.step()
        return out.loss

    graphed, graph_note = None, "disabled (--no-graph)"
    if not args.no_graph:
        try:
            from unsloth_b200.graph import GraphedTrainStep
            graphed = GraphedTrainStep(model, bucket, args.bs, args.seq, dev)
            graph_note = "fwd+bwd replayed from one CUDA graph (%d launches captured); all-reduce + AdamW outside" % graphed.launches_per_replay
        except Exception as ex:  # pragma: no cover - report and fall back to eager launches
            graphed, graph_note = None, "capture failed, eager launches: %r" % (ex,)
            torch.cuda.synchronize()

    def train_step(ids, lab):
        return graphed.step(ids, lab) if graphed is not None else eager_step(ids, lab)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(leg):
        step_fn = eager_step if leg == "profile" else train_step
        for i in range(args.warmup):
            if leg == "e2e":
                step_fn(host_ids[i], host_lab[i]).item() if graphed is not None else \
                    step_fn(host_ids[i].to(dev, non_blocking=True), host_lab[i].to(dev, non_blocking=True)).item()
            else:
                step_fn(dev_ids[i], dev_lab[i])
        barrier()
        L.launch_count = 0
        KU.GEMM_EVENTS = [] if leg == "profile" else None
        torch.cuda.reset_peak_memory_stats()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prof = leg == "resident" and os.environ.get("UB200_PROFILE_RANGE", "0") == "1"
        if prof:        # `ncu --profile-from-start off`: only the timed steps are profiled (never a bench value)
            torch.cuda.cudart().cudaProfilerStart()
        s.record()
        loss = None
        for i in range(args.warmup, total_steps):
            if leg == "e2e":
                if graphed is not None:
                    loss = step_fn(host_ids[i], host_lab[i]).item()       # H2D into the static buffers
                else:
                    loss = step_fn(host_ids[i].to(dev, non_blocking=True),
                                   host_lab[i].to(dev, non_blocking=True)).item()
            else:
                loss = step_fn(dev_ids[i], dev_lab[i])
        e.record()
        barrier()
        if prof:
            torch.cuda.cudart().cudaProfilerStop()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        ev = KU.GEMM_EVENTS
        KU.GEMM_EVENTS = None
        n_launch = L.launch_count
        if graphed is not None and leg != "profile":
            n_launch += graphed.launches_per_replay * args.steps
        return ms, n_launch, ev, (loss if isinstance(loss, float) else float(loss.item()))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res, launches, _, loss_res = timed("resident")
    ms_e2e, _, _, loss_e2e = timed("e2e")
    clocks = sampler.stop() if rank == 0 else None
    # roofline pass: the same K steps launched eagerly with a CUDA-event pair around every GEMM
    ms_prof, _, gemm_events, _ = timed("profile")
    peak_vram = torch.cuda.max_memory_allocated() / 2 ** 30

    tokens = args.bs * args.seq * world * args.steps
    value = tokens / (ms_res / 1e3)
    e2e = tokens / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM), live CUDA-event timing --------------
    peaks = {"bf16_tflops_sustained": 1400.0, "src": "fallback (B200_PROFILING.md: sustained)"}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        d = json.load(open(pp))
        peaks = {"bf16_tflops_sustained": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    roof, roof_all = None, None
    if gemm_events:
        # one entry per kernel: algorithmic flops (true LoRA rank, not the zero-padded rank block)
        # over the summed CUDA-event durations of that kernel's launches
        names = {"grouped": "ub::gemm::grouped::gemm_grouped_kernel (persistent tcgen05 cta_group::2; dense GEMMs + "
                            "rank-block producers + split-K dA/dB of a LoRA phase in one launch)",
                 "gemm2": "ub::gemm::gemm2_kernel<256, 0> (tcgen05 cta_group::2 multi-segment GEMM: q/k/v forward, gate, down, "
                          "MLP dX, lm_head / fused-CE chunks)",
                 "gemm2_glu": "ub::gemm::gemm2_kernel<256, 1> (the same GEMM with the gated activation in its epilogue: up "
                              "projection -> g, h; DW -> h, df, de in place; 16 epilogue warps)",
                 "gemm1_glu": "ub::gemm::gemm_kernel<*, 1> (single-CTA GEMM with the gated-activation epilogue)",
                 "gemm1": "ub::gemm::gemm_kernel (single-CTA tcgen05 GEMM: tails / small problems)"}
        agg = {}
        for (fl, s_, e_, info) in gemm_events:
            a_ = agg.setdefault(info["kernel"], {"flops": 0.0, "ms": 0.0, "n": 0, "rank_flops": 0.0})
            a_["flops"] += fl; a_["ms"] += s_.elapsed_time(e_); a_["n"] += 1
            a_["rank_flops"] += info.get("rank_flops", 0.0)
        traffic_by = {}
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic_by = {k: v.get("dram_bytes_per_launch") for k, v in json.load(open(tp)).get("kernels", {}).items()}
        roof_all = []
        for k_, a_ in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            ach = a_["flops"] / a_["ms"] / 1e9
            roof_all.append({"bound": "tensor", "kernel": names.get(k_, k_), "key": k_, "achieved": round(ach, 1),
                             "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": round(ach / peaks["bf16_tflops_sustained"], 3),
                             "launches_timed": a_["n"], "avg_launch_ms": round(a_["ms"] / a_["n"], 4),
                             "share_of_step": round(a_["ms"] / ms_prof, 3),
                             "rank_block_flops_share": round(a_["rank_flops"] / max(a_["flops"], 1.0), 4)})
        roof = dict(roof_all[0])
        roof.update({"traffic": traffic_by.get(roof.get("key")), "traffic_note": "ncu dram bytes of ONE representative launch "
                     "of this kernel (profiles/roofline_traffic.json), not the average over the step's shapes",
                     "peak_source": peaks["src"],
                     "flops": "algorithmic: 2*M*N*K per product with the TRUE LoRA rank for rank-block segments",
                     "timing": "CUDA-event pair around every launch during an eager pass of the same %d steps "
                               "(%.2f ms/step; events cannot be read inside a replayed graph)" % (args.steps, ms_prof / args.steps)})

    # ---- the reference's own GPU path on this box, N = 1 only ----------------
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference and not args.gradient_checkpointing:
        gpu_ref = gpu_reference_leg(model, dev_ids, dev_lab, args)
    if rank != 0:
        return 0
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, desc, threads, spread = cpu_reference_tokens_per_sec(args.model, args.cpu_sample_tokens or args.seq,
                                                                     r=args.rank, reps=3)
            cpu = {"value": round(v, 2), "unit": UNIT, "cores": threads, "kind": "reference", "sample": desc,
                   "spread_s": spread}
        except Exception as ex:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % ex}
    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "%s QLoRA NF4 r=%d bf16 seq%d bs%d/GPU (BASELINE.json configs[%d])%s" %
                                   (args.model, args.rank, args.seq, args.bs,
                                    {"cfg3": 2, "cfg5": 4}.get(args.preset, 1 if world == 1 else 3),
                                    (" sliding_window=%d" % args.sliding_window) if args.sliding_window else ""),
                       "global_batch": args.bs * world, "seq_len": args.seq, "parallelism": "dp%d" % world,
                       "layers": model.config.num_hidden_layers, "gradient_checkpointing": args.gradient_checkpointing or False, "tiled_mlp": args.tiled_mlp,
                       "optimizer": "AdamW on the flat LoRA bucket (%d params)" % bucket.numel(),
                       "cuda_graph": graph_note,
                       "keep_dequant": ("dequantised weights stay resident from a layer's forward to its backward "
                                        "(+2 B per base parameter of peak memory)") if KU.keep_dequant()
                                       else "off: weights re-dequantised in the backward",
                       "attention": "UB200_ATTENTION=%s (auto: csrc/attention.cu for window / softcap / packed rows / D=256, "
                                    "cuDNN SDPA for plain dense causal)" % os.environ.get("UB200_ATTENTION", "auto"),
                       "step_plan": os.environ.get("UB200_STEP_PLAN", "1") not in ("0", "off", "false"),
                       "glu_epilogue": "UB200_FUSED_GLU=%s (1: SwiGLU / GEGLU forward in the up projection's GEMM "
                                       "epilogue, backward in the DW GEMM's; 0: separate elementwise launches)"
                                       % os.environ.get("UB200_FUSED_GLU", "1"),
                       "l2_policy": "inputs larger than L2 (each step streams > 5 GB of NF4 weights and activations)"},
            "e2e": {"value": round(e2e, 1), "unit": UNIT, "ms_per_step": round(ms_e2e / args.steps, 2),
                    "h2d_bytes_per_step": 2 * args.bs * args.seq * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "peak_vram_gib": round(peak_vram, 2),
            "loss": {"resident_last": round(loss_res, 4), "e2e_last": round(loss_e2e, 4)},
            "roofline": roof, "roofline_by_kernel": roof_all, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "clocks": clocks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def gpu_reference_leg(model, dev_ids, dev_lab, args, attentions=("flash", "sdpa")):
    """Times the REFERENCE's GPU path on the same model object and batches: its Triton kernels
    (compiled natively for the B200), its LoRA autograd functions with their cuBLAS schedule and its
    `fast_dequantize` host path (benchmarks/ref_composite.py; the unmodified install under
    baseline/_ref).  `flash` = flash-attn 2, the backend the reference's dispatcher prefers;
    `sdpa` = cuDNN fused attention, its best case.  Returns None-with-reason if unavailable."""
    try:
        import torch
        from oracle import ref_shim
        if not ref_shim.reference_available():
            return {"unavailable": "reference not installed under baseline/_ref (DESIGN.md section 5)"}
        from benchmarks.ref_composite import time_reference_step
        res = {}
        for attn in attentions:
            r = time_reference_step(model, dev_ids, dev_lab, steps=max(3, min(args.steps, 8)),
                                    warmup=max(2, min(args.warmup, 3)), attention=attn)
            res[attn] = {"value": round(r["tokens_per_s"], 1), "unit": UNIT, "ms_per_step": round(r["ms_per_step"], 2),
                         "peak_vram_gib": round(r["peak_vram_gib"], 2), "loss_last": round(r["loss_last"], 4),
                         "steps": r["steps"], "warmup": r["warmup"]}
            torch.cuda.empty_cache()
        best = max(res.values(), key=lambda d: d["value"])
        return {"value": best["value"], "unit": UNIT, "by_attention_backend": res,
                "what": "reference Unsloth GPU path: its Triton kernels + LoRA_MLP/QKV/W autograd functions (cuBLAS "
                        "torch.matmul/addmm_ schedule) + its fast_dequantize host path on libunsloth_b200.so's "
                        "bitsandbytes-signature symbols (bitsandbytes absent), materialised-logits loss route "
                        "(llama.py:1525-1562), eager launches, torch fused AdamW; value = best backend"}
    except Exception as ex:  # pragma: no cover - report, never fail the bench line
        return {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:300])}


def run_gpu_reference(args):
    """`--impl gpu-reference`: only the reference's GPU path (same JSON shape as our line)."""
    import torch
    from unsloth_b200.patch import build_qlora_model
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    dev = torch.device("cuda", 0)
    model = build_qlora_model(args.model, r=args.rank, lora_alpha=args.rank, device=dev, seed=3407,
                              num_hidden_layers=args.layers)
    g = torch.Generator().manual_seed(1234)
    n = args.warmup + args.steps
    ids = torch.randint(0, model.config.vocab_size, (n, args.bs, args.seq), generator=g)
    lab = ids.clone()
    lab[torch.rand(lab.shape, generator=g) < 0.1] = -100
    r = gpu_reference_leg(model, ids.to(dev), lab.to(dev), args)
    line = {"impl": "gpu-reference", "metric": METRIC, "value": r.get("value"), "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "%s QLoRA NF4 r=%d bf16 seq%d bs%d/GPU" % (args.model, args.rank, args.seq, args.bs)},
            "gpu_reference": r}
    print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "gpu-reference":
        return run_gpu_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
