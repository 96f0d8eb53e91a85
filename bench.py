#!/usr/bin/env python
"""bench.py -- NeuSomaticNet candidate scoring throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]): synthetic standalone call, C=26 pileup matrices of shape
26x5x32 built from seeded random uint8[5,32,23] candidates (neusomatic_b200.synth S1) with the
bundled-layout v0.1.0 standalone checkpoint (tests/golden/std_v010.pth).  One STEP = one pass of
the hot path (NeuSomaticNet.forward + fused softmax/argmax) over one batch of --batch candidates
per GPU (default 65536: 1.09 GB of fp32 input, larger than the 126 MB L2); the default
16 steps score 1M candidates per GPU.

Printed JSON (one line, rank 0):
  value   candidates/s, whole job, inputs resident in HBM (nsc_forward_f32 on device tensors)
  e2e     candidates/s through the C-ABI host call nsc_score_host_u8 (call.py:68-83 batch step): pinned
          host uint8 pileup matrices + tag fields in (what the TSV decoder yields, dataloader.py:38-58),
          host results out, H2D/D2H inside the timed region, channel assembly on the device;
          e2e_f32 is the same through nsc_score_host_f32 (assembled fp32 [B,C,5,32] matrices in,
          PCIe-bound at 16.6 KB per candidate)
  roofline  tensor roofline of the dominant kernel family, timed live with CUDA events
  cpu_baseline  oracle/torch_port.py (the reference's forward restated on the PyTorch CPU
          operators the reference itself runs) on a bounded sample, rank 0 / N=1 only
`--impl reference` times that CPU port alone on the host cores (tier rule: the reference is
Python over PyTorch CPU kernels; nothing of ours is on that path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_CAND = {26: 2 * 62_384_240, 119: 2 * 65_241_200}      # SURVEY.md Appendix A (dense conv+fc MACs x2)
# dense MACs of the dominant kernel family (the eight 64->64 convolutions), per candidate
CONV_STACK_MACS = 5_898_240 * 2 + 16_384_000 * 2 + 2_949_120 + 8_192_000 + 1_474_560 + 4_096_000
# MACs the tensor path actually issues for that family per candidate: kernel rows in the H padding are skipped
# (13/15, 19/25, 11/15, 7/15, 13/25 of the taps), times 3 products in split mode
CONV_STACK_EXEC_MACS = 2 * (5_898_240 * 13 // 15 + 16_384_000 * 19 // 25) + 2_949_120 * 11 // 15 + 8_192_000 * 19 // 25 + \
    1_474_560 * 7 // 15 + 4_096_000 * 13 // 25
MMA_SLOTS = {0: 3, 1: 1, 3: 2}    # fp16-equivalent tcgen05 issue slots per MAC by NSC_PREC_* mode
CKPT = os.path.join(ROOT, "tests", "golden", "std_v010.pth")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "src": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_rate(C, seconds=12.0, batch=1000, threads=None):
    """Times oracle/torch_port.forward (CPU fp32, all host threads) on a bounded sample."""
    from oracle import nsnet_oracle as oracle, torch_port
    from neusomatic_b200 import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    ck = torch.load(CKPT, map_location="cpu", weights_only=True)
    p = torch_port.prepare(ck["state_dict"])
    cand = synth.uniform_pileups(batch, seed=1234, ensemble=(C == 119))
    x = torch.from_numpy(oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov,
                                                  cand.anns, int(ck["coverage_thr"])))
    torch_port.forward(p, x[:64])
    done, t0 = 0, time.perf_counter()
    while True:
        torch_port.forward(p, x)
        done += batch
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done / el, threads, "%d candidates in batches of %d over %.1f s" % (done, batch, el)


def run_reference(args, rank, world):
    if rank != 0:
        return
    C = 26
    rates = []
    for _ in range(args.warmup):
        cpu_port_rate(C, seconds=1.0, batch=args.ref_batch)
    t_budget = max(3.0, min(20.0, 120.0 / max(args.steps, 1)))
    sample = ""
    threads = os.cpu_count() or 1
    for _ in range(args.steps):
        r, threads, sample = cpu_port_rate(C, seconds=t_budget, batch=args.ref_batch)
        rates.append(r)
    v = float(np.mean(rates))
    print(json.dumps({
        "impl": "reference", "metric": "NeuSomaticNet candidates/sec", "value": v, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_budget,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic standalone call, C=26 (26x5x32), v0.1.0 standalone checkpoint; CPU sample",
                   "batch": args.ref_batch},
        "cpu_baseline": {"value": v, "unit": "candidates/s", "cores": threads, "kind": "port",
                         "sample": "oracle/torch_port.py (PyTorch CPU operators, fp32): " + sample},
        "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_train(args, rank, local_rank, world, dev):
    """BASELINE.json configs[4] (not the headline metric): one train.py optimisation step per bench step on a synthetic
    batch per GPU; gradients summed over ranks with one NCCL all-reduce of the flat 873 385-float vector."""
    import torch.distributed as dist
    from neusomatic_b200 import synth
    from neusomatic_b200.train import Trainer
    from oracle import nsnet_oracle as oracle     # input assembly of the synthetic batch only (outside the timed region)
    B = min(args.batch, 1024)
    cand = synth.structured_pileups(B, seed=100 + rank)
    x = torch.from_numpy(oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, None, 100)).to(dev)
    vt = torch.tensor([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], dtype=torch.int32, device=dev)
    ln = torch.tensor([min(int(t.split(".")[6]), 3) for t in cand.tags], dtype=torch.int32, device=dev)
    ce = torch.from_numpy(cand.center).float().to(dev)
    w1 = torch.tensor([0.2, 0.21, 0.13, 0.16], device=dev)
    w3 = torch.tensor([0.12, 0.2, 0.23, 0.24], device=dev)
    tr = Trainer(synth.random_state_dict(26, seed=11), 26, device=local_rank, max_batch=B)
    for _ in range(args.warmup):
        tr.step(x, vt, ce, ln, w1, w3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = tr.step(x, vt, ce, ln, w1, w3)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "NeuSomaticNet train samples/sec (fwd+bwd+SGD)", "value": world * B * args.steps / (ms / 1000.0),
                          "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "split fp16 x3 on tcgen05 with fp32 accumulate for the 64->64 convolutions (forward, data and weight "
                                   "gradient); f32 elsewhere (NSC_TRAIN_FP32=1: all f32)", "data": "synthetic",
                          "config": {"workload": "train.py step (BASELINE.json configs[4]): C=26, SGD lr 0.01 momentum 0.9, "
                                                 "BatchNorm batch statistics per GPU, one NCCL all-reduce of 873385 floats per step",
                                     "batch_per_gpu": B}, "loss": float(losses[0])}), flush=True)
    tr.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="candidates per GPU per step")
    ap.add_argument("--precision", type=int, default=None, help="NSC_PREC_* (default: library default)")
    ap.add_argument("--ref-batch", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--train", action="store_true", help="BASELINE configs[4]: train.py step (fwd+bwd+SGD), data-parallel with one NCCL all-reduce")
    ap.add_argument("--ensemble", action="store_true", help="BASELINE configs[3]: C=119 ensemble model (not the default metric config)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from neusomatic_b200 import Engine, synth
    from neusomatic_b200.call import load_checkpoint

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.train:
        run_train(args, rank, local_rank, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return

    C = 119 if args.ensemble else 26
    sd, meta = load_checkpoint(os.path.join(ROOT, "tests", "golden", "ens_v010.pth") if args.ensemble else CKPT)
    eng = Engine(sd, C, device=local_rank, precision=args.precision)
    B = args.batch

    # synthetic S1 candidates, seed = 1234 + rank; a 4096-candidate pool tiled to B
    from neusomatic_b200 import anns_to_255
    pool = synth.uniform_pileups(4096, seed=1234 + rank, ensemble=args.ensemble)
    reps = (B + 4095) // 4096
    mat_h = torch.from_numpy(np.tile(pool.matrices, (reps, 1, 1, 1))[:B]).pin_memory()
    meta_h = torch.from_numpy(np.tile(pool.meta, (reps, 1))[:B].copy()).pin_memory()
    ann_h = torch.from_numpy(np.tile(anns_to_255(pool.anns), (reps, 1))[:B].copy()).pin_memory() if args.ensemble else None
    mat_d, meta_d = mat_h.to(dev), meta_h.to(dev)
    ann_d = ann_h.to(dev) if args.ensemble else None
    x_d = eng.assemble(mat_d, meta_d, ann_d, float(meta["coverage_thr"]))        # fp32 [B,C,5,32] in HBM
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident inputs, kernel-only path ---------------------------------------
    for _ in range(args.warmup):
        eng.forward_f32(x_d, extras=True)
    barrier()
    eng.get_profile()
    eng.set_profile(True)
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        outs = eng.forward_f32(x_d, extras=True)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    prof = eng.get_profile()
    eng.set_profile(False)
    launches = eng.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1000.0)

    # ---- e2e: host buffers through the C ABI -------------------------------------------------------
    e2e = e2e_u8 = None
    if not args.no_e2e:
        x_h = torch.empty((B, C, 5, 32), dtype=torch.float32).pin_memory()
        x_h.copy_(x_d)
        host_out = eng._host_out(B, True, pin=True)
        k = max(2, min(args.steps, 6))
        res = {}
        for name in ("f32", "u8"):
            fn = (lambda: eng.score_host_f32(x_h, out=host_out)) if name == "f32" else \
                 (lambda: eng.score_host_u8(mat_h, meta_h, ann_h, float(meta["coverage_thr"]), out=host_out))
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(k):
                fn()     # synchronises: results are in host memory when it returns
            torch.cuda.synchronize()
            el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            res[name] = world * B * k / float(el.item())
        d2h = B * 19 * 4
        e2e = {"value": res["u8"], "unit": "candidates/s", "h2d_bytes_per_step": B * (3680 + 12 + (372 if args.ensemble else 0)),
               "d2h_bytes_per_step": d2h, "api": "nsc_score_host_u8 (pinned uint8 [B,5,32,23] + int32 [B,3] in, host results out)"}
        e2e_u8 = {"value": res["f32"], "unit": "candidates/s", "h2d_bytes_per_step": B * C * 160 * 4,
                  "d2h_bytes_per_step": d2h, "api": "nsc_score_host_f32 (pinned fp32 [B,C,5,32] in, host results out)"}

    # ---- roofline of the dominant kernel family (the eight 64->64 convolutions) -------------------
    pk = peaks()
    conv_ms = sum(v[0] for k_, v in prof.items() if k_.startswith("res"))
    conv_launches = sum(v[1] for k_, v in prof.items() if k_.startswith("res"))
    total_ms = sum(v[0] for v in prof.values())
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]     # ncu capture, 32768-candidate launches
    if conv_ms > 0:
        cands = B * args.steps
        achieved = cands * 2 * CONV_STACK_MACS / (conv_ms / 1000.0) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained",
                "kernel": eng.kernel_name, "kernel_launches": conv_launches,
                "kernel_ms_per_launch": conv_ms / max(conv_launches, 1),
                "kernel_share_of_step": conv_ms / total_ms if total_ms else None,
                "algorithmic_flops_per_candidate": 2 * CONV_STACK_MACS,
                # tensor-pipe occupancy in fp16-instruction equivalents: 3 per MAC (split-f16), 2 per MAC (split8: one fp16
                # + two fp8 products, an fp8 instruction covers twice the K of an fp16 one in the same issue slot)
                "executed_tensor_tflops": cands * 2 * CONV_STACK_EXEC_MACS * MMA_SLOTS.get(eng.precision, 1) /
                                          (conv_ms / 1000.0) / 1e12 if eng.precision != 2 else None,
                "executed_frac_of_peak": (cands * 2 * CONV_STACK_EXEC_MACS * MMA_SLOTS.get(eng.precision, 1) /
                                          (conv_ms / 1000.0) / 1e12 / pk["bf16_sustained"]) if eng.precision != 2 else None,
                "whole_net_frac": (value / world) * FLOPS_PER_CAND[C] / 1e12 / pk["bf16_sustained"],
                "per_kernel_ms": {k_: round(v[0] / args.steps, 4) for k_, v in prof.items()}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, threads, sample = cpu_port_rate(C, seconds=12.0, batch=args.ref_batch)
        cpu = {"value": r, "unit": "candidates/s", "cores": threads, "kind": "port",
               "sample": "oracle/torch_port.py (PyTorch CPU operators, fp32): " + sample}

    if rank == 0:
        prec_names = {0: "split-f16 (fp16 hi+lo operands, 3 tcgen05 MMAs per MAC, fp32 accumulate)", 1: "f16", 2: "f32",
                      3: "f16 + fp8 corrections (fp16 main product, two e4m3 correction products scaled by 2^15, fp32 accumulate)"}
        print(json.dumps({
            "metric": "NeuSomaticNet candidates/sec", "value": value, "unit": "candidates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec_names.get(eng.precision, str(eng.precision)), "data": "synthetic",
            "config": {"workload": ("synthetic ensemble call (BASELINE.json configs[3]): C=119, 119x5x32 matrices, v0.1.0 ensemble "
                                    "checkpoint, S1 uniform pileups") if args.ensemble else
                                   ("synthetic standalone call (BASELINE.json configs[2]): C=26, 26x5x32 matrices, "
                                    "v0.1.0 standalone checkpoint, S1 uniform pileups"),
                       "batch_per_gpu": B, "candidates_per_step": world * B,
                       "l2_policy": "inputs larger than L2 (%.2f GB fp32 per step per GPU)" % (B * C * 640 / 1e9),
                       "sharding": "contiguous candidate ranges per rank, no collective on the data path"},
            "e2e": e2e, "e2e_f32": e2e_u8, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "clocks": clocks,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
