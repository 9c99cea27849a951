#!/usr/bin/env python
"""Benchmark of the SDUMC hot path (BASELINE.json: train samples/sec, fwd + bwd + distill + Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

N = 1 runs BASELINE config 2: full + text-missing self-distillation train step, bf16 storage / fp32
accumulate, batch 512 per GPU, synthetic MER2024-shaped features "S0" (dims 1024/4096/1024/4096, frames
384/64/256/64).  N > 1 (launched by torch.distributed.run) is data parallel, 512 samples per GPU (weak
scaling), global RnC / RMSE semantics kept by the exchanges in sdumc_b200/trainer.py.

`--impl reference` times the reference algorithm on the host CPU cores (the oracle port of the reference
modules: /root/reference is not present on the GPU box) on a bounded sample of the same workload.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "train_samples_per_sec"
UNIT = "samples/s"
F_TRAIN_PER_SAMPLE = 2.408e9      # algorithmic FLOPs / sample of the S0 train step
F_INFER_PER_SAMPLE = 1.004e9      # ... of the two-pass scoring forward


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, read from this round's `ncu --set full` captures
    (profiles/r2_traffic.json, written by tools/ncu_summary.py from the .ncu-rep files); {} until captured."""
    p = ROOT / "profiles" / "r2_traffic.json"
    try:
        return json.loads(p.read_text())
    except Exception:  # noqa: BLE001
        return {}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference train step on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_steps(n_steps: int, warmup: int, B: int = 32):
    import torch
    from oracle import sdumc_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    P = O.init_params(O.S0_DIMS, seed=100, gain=1.0)
    b = O.synth_batch(B, O.S0_DIMS, O.S0_FRAMES, seed=1234)
    sites = O.dropout_sites()
    ps = dict(sites)
    st = {}

    def make_drop():
        def drop(site, x):   # train mode: torch dropout, as the reference's nn.Dropout modules
            return torch.nn.functional.dropout(x, ps[site], True)
        return drop

    times = []
    for i in range(warmup + n_steps):
        t0 = time.perf_counter()
        O.train_step(P, st, b["audio"], b["text"], b["feat4"], b["video"], b["vals"], lr=1e-4, weight_decay=1e-5,
                     drop0=make_drop(), drop1=make_drop())
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return B, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, times = cpu_reference_steps(args.steps, args.warmup, B=32)
    total = sum(times)
    value = B * len(times) / total
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "SDUMC full + text-missing self-distill train step, S0 features "
                               "(dims 1024/4096/1024/4096, frames 384/64/256/64)", "batch_per_step": B,
                   "note": "reference algorithm on the host CPU (oracle port of the reference modules), torch eager "
                           "fp32, bounded sample: batch 32 per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} train steps of batch {B} (S0 shapes), torch fp32, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sdumc_b200 import _lib, ops
    from sdumc_b200.data import S0_DIMS, S0_FRAMES, pin, synth_batch
    from sdumc_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sdumc_b200 has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    B = args.batch
    peaks = load_peaks()

    tr = Trainer(S0_DIMS, B, S0_FRAMES, dev, lr=1e-4, weight_decay=1e-5, seed=100, process_group=pg)
    batch = synth_batch(B, S0_DIMS, S0_FRAMES, seed=1234 + rank, device=dev)
    host = pin(batch)
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("audio", "text", "video", "feat4", "vals"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def trace(msg):
        if os.environ.get("SDUMC_BENCH_TRACE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident inputs: whole-step throughput ----
    tr.load_batch(batch["audio"], batch["text"], batch["video"], batch["feat4"], batch["vals"])
    _lib.KERNEL_LAUNCHES[0] = 0
    tr.train_step()                       # eager step: counts our kernel launches per step
    launches = _lib.KERNEL_LAUNCHES[0]
    for _ in range(max(args.warmup, 3) - 1):
        tr.train_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            tr.train_step()
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    trace("device-resident timing done")
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    terms = tr.terms.tolist()

    # ---- end to end: pinned host batch -> H2D -> step -> D2H of the loss, every step ----
    # The public API double-buffers host batches: stage_batch() starts the H2D copy of the next batch on a copy
    # stream while the current step runs, commit_staged() makes it current.  Every timed step still pays its own
    # 1.2 GB H2D copy (n_e2e copies inside the region) and a synchronous D2H read of its loss.
    hb = (host["audio"], host["text"], host["video"], host["feat4"], host["vals"])

    def e2e_steps(n):
        tr.stage_batch(*hb)
        for i in range(n):
            tr.commit_staged()
            if i + 1 < n:
                tr.stage_batch(*hb)
            tr.train_step()
            float(tr.terms[6].item())

    e2e_steps(2)
    barrier()
    n_e2e = max(3, min(args.steps, 10))
    e0.record()
    e2e_steps(n_e2e)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    trace("e2e timing done")
    e2e_value = world * B * n_e2e / (ms_e2e / 1e3)

    # ---- scoring path (both eval passes, main_frame_val_text_missing_inference.py:158-175), device-resident ----
    tr.load_batch(batch["audio"], batch["text"], batch["video"], batch["feat4"], batch["vals"])
    for _ in range(3):
        tr.score()
    barrier()
    e0.record()
    for _ in range(args.steps):
        tr.score()
    e1.record()
    barrier()
    ms_score = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    trace("scoring timing done")

    # ---- scoring at BASELINE config 5's batch (B = 128, shell/main_text_missing_icassp_inference.sh:5) ----
    Bs = 128
    tr.load_batch(*(batch[k][:Bs] for k in ("audio", "text", "video", "feat4")), batch["vals"][:Bs])
    for _ in range(3):
        tr.score()
    barrier()
    e0.record()
    for _ in range(args.steps):
        tr.score()
    e1.record()
    barrier()
    ms_score128 = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    tr.load_batch(batch["audio"], batch["text"], batch["video"], batch["feat4"], batch["vals"])

    # ---- BASELINE config 5 end to end: 100,000 utterances scored in reference batches of 128, sharded over the ranks at
    #      whole-batch granularity, every batch gathered from the HBM-resident feature store by the collate kernel,
    #      both eval passes replayed as a CUDA graph, predictions + all 8 embeddings copied to pinned host memory.
    #      The store holds 2048 distinct synthetic utterances (4.8 GB; 100k distinct ones would be 236 GB), visited
    #      cyclically: every one of the 100k utterances is gathered and scored in full. ----
    inf = None
    if not args.no_inference:
        from sdumc_b200.dataset import DeviceStore4F, batch_chunks
        n_total, bs_inf, pool = 100_000, 128, 2048
        pb = synth_batch(pool, S0_DIMS, S0_FRAMES, seed=4242, device=dev)
        store = DeviceStore4F.from_device_batch(pb, pb["vals"])
        del pb
        mine = batch_chunks(n_total, bs_inf, rank, world)
        n_mine = sum(len(c) for c in mine)
        shapes = {"val_preds_full": (1,), "val_preds_missing": (1,), "full_rep": (128,), "missing_rep": (128,),
                  "full_rnc": (64,), "missing_rnc": (64,), "text_rep_query_full": (256,), "text_rep_query_missing": (256,),
                  "text_rep_full": (7, 128), "text_rep_missing": (7, 128)}
        import math
        width = sum(math.prod(sh) for sh in shapes.values())
        host_pack = torch.empty(n_mine, width, dtype=torch.float32).pin_memory()      # one packed row per utterance

        def infer_pass():
            lo = 0
            for c in mine:
                tr.load_from_store(store, [i % pool for i in c], labels=False)
                tr.score()
                host_pack[lo:lo + len(c)].copy_(tr.last_packed, non_blocking=True)   # all 10 outputs: one D2H copy
                lo += len(c)
        for c in mine[:3] + mine[-1:]:              # warm-up: capture the graphs of both batch shapes
            tr.load_from_store(store, [i % pool for i in c])
            tr.score()
            tr.load_from_store(store, [i % pool for i in c])
            tr.score()
        barrier()
        e0.record()
        infer_pass()
        e1.record()
        barrier()
        ms_inf = max_over_ranks(e0.elapsed_time(e1))
        host_out = tr.unpack_scores(host_pack)
        assert set(host_out) == set(shapes) and all(tuple(host_out[k].shape[1:]) == shapes[k] for k in shapes)
        d2h = host_pack.numel() * 4
        inf = {"value": n_total / (ms_inf * 1e-3), "unit": "utterances/s", "utterances": n_total, "batch": bs_inf,
               "ms_total": ms_inf, "batches_per_rank": len(mine), "d2h_bytes_per_rank": d2h,
               "finite": bool(torch.isfinite(host_out["val_preds_full"]).all()),
               "roofline": {"bound": "tensor", "achieved": n_total / world / (ms_inf * 1e-3) * F_INFER_PER_SAMPLE / 1e12,
                            "peak": peaks["tc_sustained"], "unit": "TFLOP/s",
                            "frac": n_total / world / (ms_inf * 1e-3) * F_INFER_PER_SAMPLE / 1e12 / peaks["tc_sustained"]},
               "note": "main_frame_val_text_missing_inference path: collate from the device store + two eval passes "
                       "(graph replay) + D2H of predictions and embeddings, per batch of 128; whole job over all ranks"}
        del store, host_out, host_pack
        tr.load_batch(batch["audio"], batch["text"], batch["video"], batch["feat4"], batch["vals"])

    # ---- roofline: the step is bounded by the tensor cores (99 % of its FLOPs are dense
    #      contractions; ideal-fusion intensity ~1000 FLOP/B vs a ridge of ~213); the headline fraction is the whole
    #      step's algorithmic FLOPs over the SUSTAINED measured bf16 peak.  Beside it, the dominant kernels timed alone
    #      (L2 flushed between launches, CUDA events on the launching stream) against the burst peaks, each with §8d's
    #      algorithmic bytes (inputs once + outputs once) and, separately, the bytes this design really moves. ----
    traffic = measured_traffic()
    step_tf = value / world * F_TRAIN_PER_SAMPLE / 1e12
    roof = {"bound": "tensor", "kernel": "whole train step (all kernels of one CUDA-graph replay)",
            "achieved": step_tf, "peak": peaks["tc_sustained"], "unit": "TFLOP/s", "frac": step_tf / peaks["tc_sustained"],
            "traffic": traffic.get("step"), "algorithmic_flops_per_sample": F_TRAIN_PER_SAMPLE,
            "peak_source": peaks["src"] + " (sustained: measured inside a long step)"}
    roof_extra = []
    if rank == 0:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def time_alone(go, n=10):
            for _ in range(3):
                go()
            ts = []
            for _ in range(n):
                flush.zero_()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                go()
                b_.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            return sum(ts) / len(ts)

        def entry(kernel, ms, flops, alg_bytes, design_bytes, key):
            tf, gbs = flops / (ms * 1e-3) / 1e12, alg_bytes / (ms * 1e-3) / 1e9
            return {"kernel": kernel, "ms_per_launch": ms, "bound": "tensor" if flops / alg_bytes > peaks["tc_burst"] * 1e3 / peaks["hbm"] else "hbm",
                    "tensor_tflops": tf, "tensor_frac": tf / peaks["tc_burst"], "algorithmic_bytes": alg_bytes,
                    "hbm_gbs": gbs, "hbm_frac": gbs / peaks["hbm"], "design_bytes": design_bytes,
                    "design_hbm_frac": design_bytes / (ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": traffic.get(key),
                    "peak_source": peaks["src"] + " (burst: kernel timed alone)"}

        La, Da = S0_FRAMES[0], S0_DIMS[0]
        rows = B * La
        X = tr.inputs["a"].view(rows, Da)
        Wt = tr.W.bf16("frame_dim_reshape_0.weight")
        bias = tr.W.f32("frame_dim_reshape_0.bias")
        for extra in bench_kernel_probes(tr, ops, torch, dev, B, La, Da, rows, X, Wt, bias, time_alone, entry):
            roof_extra.append(extra)

    # ---- BASELINE config 4 "S1": attention-heavy stress - 256 frames per modality, general_dim 1024, batch 1024 per run
    #      split over the GPUs (strong scaling of a fixed job) ----
    s1 = None
    if not args.no_stress:
        F_S1 = 30.77e9                                  # algorithmic FLOPs / sample
        Bs1 = 1024 // world
        tr.close()
        tr1 = Trainer(S0_DIMS, Bs1, (256, 256, 256, 256), dev, lr=1e-4, weight_decay=1e-5, seed=100, process_group=pg,
                      general_dim=1024)
        b1 = synth_batch(Bs1, S0_DIMS, (256, 256, 256, 256), seed=777 + rank, device=dev)
        tr1.load_batch(b1["audio"], b1["text"], b1["video"], b1["feat4"], b1["vals"])
        for _ in range(3):
            tr1.train_step()
        barrier()
        n1 = max(3, args.steps // 2)
        e0.record()
        for _ in range(n1):
            tr1.train_step()
        e1.record()
        barrier()
        ms1 = max_over_ranks(e0.elapsed_time(e1)) / n1
        v1 = 1024 / (ms1 * 1e-3)
        s1 = {"value": v1, "unit": "samples/s", "ms_per_step": ms1, "global_batch": 1024, "batch_per_gpu": Bs1,
              "general_dim": 1024, "frames": [256, 256, 256, 256], "scaling": "strong", "steps": n1,
              "loss_total": float(tr1.terms[6].item()),
              "roofline": {"bound": "tensor", "achieved": v1 / world * F_S1 / 1e12, "peak": peaks["tc_sustained"],
                           "unit": "TFLOP/s", "frac": v1 / world * F_S1 / 1e12 / peaks["tc_sustained"]}}
        tr1.close()
        del tr1, b1

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc, times = cpu_reference_steps(n_steps=8, warmup=1, B=32)
        cores = os.cpu_count() or 1
        cpu = {"value": Bc * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{len(times)} train steps of batch {Bc} (S0 shapes) of the oracle port, torch fp32, "
                         f"{cores} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "SDUMC full + text-missing self-distill train step (fwd + bwd + 6-term loss + Adam), "
                                   "S0 features dims 1024/4096/1024/4096, frames 384/64/256/64",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "inputs (1.2 GB bf16 per GPU) larger than L2; no explicit flush",
                       "cuda_graph": bool(tr.use_graph)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": n_e2e},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": clk.summary(),
            "scoring": {"value": world * B / (ms_score * 1e-3), "unit": "samples/s", "ms_per_batch": ms_score,
                        "note": "two eval passes per sample, CUDA-graph replay, inputs resident",
                        "roofline": {"bound": "tensor", "achieved": B / (ms_score * 1e-3) * F_INFER_PER_SAMPLE / 1e12,
                                     "peak": peaks["tc_sustained"], "unit": "TFLOP/s",
                                     "frac": B / (ms_score * 1e-3) * F_INFER_PER_SAMPLE / 1e12 / peaks["tc_sustained"]}},
            "scoring_b128": {"value": world * Bs / (ms_score128 * 1e-3), "unit": "samples/s", "ms_per_batch": ms_score128,
                             "note": "BASELINE config 5 batch size (128 utterances per batch and GPU), two eval passes, "
                                     "CUDA-graph replay, inputs resident",
                             "roofline": {"bound": "tensor", "achieved": Bs / (ms_score128 * 1e-3) * F_INFER_PER_SAMPLE / 1e12,
                                          "peak": peaks["tc_sustained"], "unit": "TFLOP/s",
                                          "frac": Bs / (ms_score128 * 1e-3) * F_INFER_PER_SAMPLE / 1e12 / peaks["tc_sustained"]}},
            "inference_100k": inf,
            "stress_s1": s1,
            "roofline": roof,
            "roofline_extra": roof_extra,
            "cpu_baseline": cpu,
            "loss_terms": dict(zip(("mse_full", "mse_missing", "rmse_text_hidden", "rmse_cross_text", "rmse_fused",
                                    "rnc", "total"), terms[:7])),
        }
        print(json.dumps(line), flush=True)
    tr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_kernel_probes(tr, ops, torch, dev, B, La, Da, rows, X, Wt, bias, time_alone, entry):
    """The frame-level kernels of the audio stream (196,608 rows at B = 512) as the train step runs them."""
    G = 256
    h = rows * G * 2                                            # one bf16 [rows,256] tensor
    out = []
    # (1) in-projection X W^T + b with the frame-dropout copies of the 2 blocks x 2 passes that consume it.
    #     §8d algorithmic bytes: X once + W + H once; design bytes: the copies actually written.
    tg = [torch.empty(rows, G, dtype=torch.bfloat16, device=dev) for _ in range(4)]
    ms = time_alone(lambda: ops.gemm(X, Wt, M=rows, N=G, K=Da, bias=bias, epi_kind=ops.EPI_INPROJ, targets=tg,
                                     target_sites=[1, 2, 3, 4], seed=5, step=1))
    out.append(entry("gemm_tcgen05_kernel<256,bf16,INPROJ> (audio in-projection, train mode)", ms, 2.0 * rows * G * Da,
                     X.numel() * 2 + Wt.numel() * 2 + h, X.numel() * 2 + Wt.numel() * 2 + 4 * h, "inproj"))
    del tg
    # (2) the K = 256 frame GEMM family (33 % of the round-1 step): Cross_Attention key projection tanh(X' W^T + b)
    Xp = torch.randn(rows, G, device=dev).bfloat16()
    Wk = tr.W.bf16("cross_att_fra2utt_0.input_proj.weight")
    bk = tr.W.f32("cross_att_fra2utt_0.input_proj.bias")
    Kt = torch.empty(rows, G, dtype=torch.bfloat16, device=dev)
    ms = time_alone(lambda: ops.gemm(Xp, Wk, M=rows, N=G, K=G, bias=bk, act=ops.ACT_TANH, out_bf16=Kt))
    out.append(entry("gemm_tcgen05_kernel<256,bf16,GENERIC> (audio key projection, K = 256)", ms, 2.0 * rows * G * G,
                     2 * h + Wk.numel() * 2, 2 * h + Wk.numel() * 2, "keyproj"))
    # (3) attention backward, 7 queries: reads X' and K, writes dZ and dH
    Kt.copy_(torch.tanh(torch.randn(rows, G, device=dev)).bfloat16())
    Pm = torch.softmax(torch.randn(B, La, 7, device=dev), dim=1).reshape(rows, 7).contiguous()
    dOut, Opre = torch.randn(B, 7, G, device=dev), torch.randn(B, 7, G, device=dev)
    Qp = torch.randn(B, 7, G, device=dev) * 0.5
    dZ = torch.empty(rows, G, dtype=torch.bfloat16, device=dev)
    dH = torch.empty(rows, G, dtype=torch.bfloat16, device=dev)
    dQp, db = torch.zeros(B, 7, G, device=dev), torch.zeros(G, device=dev)
    ms = time_alone(lambda: ops.attn_bwd(Xp, Kt, Pm, dOut, dout_stride_b=7 * G, O_pre=Opre, Qp=Qp, qp_stride_b=7 * G, B=B,
                                         L=La, nq=7, out_drop_p=0.5, out_site=3, dZ=dZ, dH=dH, dh_mode=0, fmask_site=2,
                                         dQp=dQp, dqp_stride_b=7 * G, db=db, seed=5, step=1))
    nb = 4 * h + Pm.numel() * 4 + 4 * B * 7 * G * 4
    out.append(entry("attn_bwd_kernel<7> (audio Cross_Attention backward, row-wise part)", ms, 2.0 * rows * 4 * 8 * G, nb, nb,
                     "attn_bwd"))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", dest="no_inference", action="store_true", help="skip the config-5 inference leg")
    ap.add_argument("--no-stress", dest="no_stress", action="store_true", help="skip the config-4 (G = 1024) stress leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
