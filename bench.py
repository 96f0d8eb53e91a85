#!/usr/bin/env python
"""bench.py -- NeuSomaticNet candidate scoring throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]): synthetic standalone call, C=26 pileup matrices of shape 26x5x32 built from seeded
random uint8[5,32,23] candidates (neusomatic_b200.synth S1) with the bundled-layout v0.1.0 standalone checkpoint
(tests/golden/std_v010.pth).  One STEP = one pass of the hot path (NeuSomaticNet.forward + fused softmax/argmax) over one
batch of --batch candidates per GPU (default 65536: 1.09 GB of fp32 input, larger than the 126 MB L2); the default 16 steps
score 1M candidates per GPU.

Timing: a REGION = exactly K steps between two CUDA events on the launching stream, a barrier + synchronize on both
sides.  A region of the default size lasts ~0.2 s, too short to tell a hiccup of one rank from the steady state, so the
region is repeated until >= 2 s have been timed (the same number of repeats on every rank); the line reports the MEDIAN
region (`ms_per_step` = its max over ranks / K), every region (`region_ms`) and the per-rank times of the median region
(`ms_per_rank`).

Printed JSON (one line, rank 0):
  value     candidates/s, whole job, inputs resident in HBM (nsc_forward_f32 on device tensors)
  e2e       candidates/s through the C-ABI host call nsc_score_host_u8 (call.py:68-83 batch step): pinned host uint8 pileup
            matrices + tag fields in (what the TSV decoder yields, dataloader.py:38-58), host results out, H2D/D2H inside
            the timed region, channel assembly on the device; e2e_f32: the same through nsc_score_host_f32
  roofline  tensor roofline of the dominant kernel family (the eight 64->64 convolutions), timed live with CUDA events,
            against BOTH measured bf16 peaks (sustained = `frac`, burst = `frac_burst`); `cublas_bf16_same_box` = what an
            8192^3 bf16 cuBLAS GEMM sustains on this box right after the timed regions (informational: the boxes of the pool
            hold different clocks under the power cap), `executed_vs_cublas_same_box` = executed tensor instructions vs that
  cpu_baseline  oracle/torch_port.py (the reference's forward restated on the PyTorch CPU operators the reference itself
            runs) on a bounded sample, rank 0 / N=1 only
  extra     records measured in the same run with the same build: `ensemble` (BASELINE configs[3], C=119: value + e2e),
            `train` (configs[4]: one train.py step per GPU + NCCL all-reduce), `gpu_incumbent` (the stock nn.Module in
            eager PyTorch on the same GPU: fp32 and TF32-allowed, inference and fwd+bwd+SGD -- informational, N=1 only),
            `e2e_tsv` (a reference-format candidate TSV -> dictionaries + none.vcf filter through the stage / device-inflate +
            score / file pipeline, with the same call on the host-inflate path and the oracle's chain beside it),
            `e2e_tsv_ensemble` (the same for C=119 lines with 93 annotation fields), `whole_call` (call_neusomatic end to end:
            TSVs + FASTA + checkpoint -> pred.vcf + none.vcf) -- these three N=1 only
`--impl reference` times the CPU port alone on the host cores (tier rule: the reference is Python over PyTorch CPU
kernels and has no setup.py to install; nothing of ours is on that path).
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
# (13/15, 19/25, 11/15, 7/15, 13/25 of the taps)
CONV_STACK_EXEC_MACS = 2 * (5_898_240 * 13 // 15 + 16_384_000 * 19 // 25) + 2_949_120 * 11 // 15 + 8_192_000 * 19 // 25 + \
    1_474_560 * 7 // 15 + 4_096_000 * 13 // 25
MMA_SLOTS = {0: 3, 1: 1, 3: 2}    # fp16-equivalent tcgen05 issue slots per MAC by NSC_PREC_* mode
CKPT = os.path.join(ROOT, "tests", "golden", "std_v010.pth")
CKPT_ENS = os.path.join(ROOT, "tests", "golden", "ens_v010.pth")
MIN_TIMED_MS = 2000.0             # repeat the K-step region until this much has been timed
PREC_NAMES = {0: "split-f16 (fp16 hi+lo operands, 3 tcgen05 MMAs per MAC, fp32 accumulate, round-toward-zero compensated)", 1: "f16", 2: "f32",
              3: "f16 + fp8 corrections (fp16 main product, two e4m3 correction products scaled by 2^15, fp32 accumulate)"}


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that touch oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(C, seconds=12.0, batch=1000, threads=None):
    """Times oracle/torch_port.forward (CPU fp32, all host threads) on a bounded sample."""
    from oracle import nsnet_oracle as oracle, torch_port
    from neusomatic_b200 import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    ck = torch.load(CKPT if C == 26 else CKPT_ENS, map_location="cpu", weights_only=True)
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
    return done / el, threads, "%d candidates in batches of %d over %.1f s" % (done, batch, el), el


def cpu_tsv_rate(path, C, seconds=8.0, batch=1000):
    """The reference's file -> dictionaries chain restated on the CPU (oracle: TSV line decode, channel assembly, forward,
    call_variants loop) over the first candidates of `path` for a bounded time: the CPU number beside `e2e_tsv`."""
    from oracle import nsnet_oracle as oracle, torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    ck = torch.load(CKPT if C == 26 else CKPT_ENS, map_location="cpu", weights_only=True)
    p = torch_port.prepare(ck["state_dict"])
    done, t0 = 0, time.perf_counter()
    with open(path) as f:
        while time.perf_counter() - t0 < seconds:
            lines = [f.readline() for _ in range(batch)]
            lines = [l for l in lines if l]
            if not lines:
                break
            dec = [oracle.decode_tsv_line(l) for l in lines]
            tags = [d[0] for d in dec]
            mats = np.stack([d[1] for d in dec])
            tagf = [oracle.parse_tag(t) for t in tags]
            anns = np.array([d[2] for d in dec], np.float32) if C == 119 else None
            x = oracle.assemble_channels(mats, np.array([t[1] for t in tagf]), np.array([t[3] for t in tagf]),
                                         np.array([t[4] for t in tagf]), anns, int(ck["coverage_thr"]))
            o1, o2, o3 = torch_port.forward(p, torch.from_numpy(x))
            oracle.call_variants(np.asarray(o1), np.asarray(o2), np.asarray(o3), tags)
            done += len(lines)
    el = time.perf_counter() - t0
    return done / el, "%d candidates in batches of %d over %.1f s" % (done, batch, el)


def workload_config(args, world):
    """The `config` object of the JSON line; identical for both arms (the reference arm times a bounded sample of it)."""
    C = 119 if getattr(args, "ensemble", False) else 26
    B = args.batch
    return {"workload": ("synthetic ensemble call (BASELINE.json configs[3]): C=119, 119x5x32 matrices, v0.1.0 ensemble "
                         "checkpoint, S1 uniform pileups") if C == 119 else
                        ("synthetic standalone call (BASELINE.json configs[2]): C=26, 26x5x32 matrices, "
                         "v0.1.0 standalone checkpoint, S1 uniform pileups"),
            "batch_per_gpu": B, "candidates_per_step": world * B,
            "l2_policy": "inputs larger than L2 (%.2f GB fp32 per step per GPU)" % (B * C * 640 / 1e9),
            "sharding": "contiguous candidate ranges per rank, no collective on the data path"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    C = 119 if args.ensemble else 26
    rates, secs = [], []
    t_budget_w = args.ref_seconds if args.ref_seconds else 1.0
    for _ in range(args.warmup):
        cpu_port_rate(C, seconds=min(1.0, t_budget_w), batch=args.ref_batch)
    t_budget = args.ref_seconds if args.ref_seconds else max(3.0, min(20.0, 120.0 / max(args.steps, 1)))
    sample = ""
    threads = os.cpu_count() or 1
    for _ in range(args.steps):
        r, threads, sample, el = cpu_port_rate(C, seconds=t_budget, batch=args.ref_batch)
        rates.append(r)
        secs.append(el)
    v = float(np.mean(rates))
    print(json.dumps({
        "impl": "reference", "metric": "NeuSomaticNet candidates/sec", "value": v, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1)),
        "timing": "each step = a bounded CPU sample of the workload, scored %d candidates at a time" % args.ref_batch,
        "cpu_baseline": {"value": v, "unit": "candidates/s", "cores": threads, "kind": "port",
                         "sample": "oracle/torch_port.py (PyTorch CPU operators, fp32): " + sample},
        "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# timing helpers
# ---------------------------------------------------------------------------------------------------------------------
class Ranks:
    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def gather(self, values):
        """[len(values)] per rank -> [world, len(values)] on every rank."""
        t = torch.tensor(values, dtype=torch.float64, device=self.dev).reshape(1, -1)
        if self.world == 1:
            return t.cpu().numpy()
        import torch.distributed as dist
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return torch.cat(out, 0).cpu().numpy()


def timed_regions(rk, step_fn, steps, warmup, device_events=True, min_ms=MIN_TIMED_MS, max_repeats=40):
    """Runs `warmup` untimed steps, then K-step regions (barrier + synchronize on both sides, CUDA events on the current
    stream or the host clock for calls that synchronise themselves) until >= min_ms have been timed.  Returns
    (median region: max over ranks in ms, [region max over ranks], per-rank ms of the median region)."""
    for _ in range(warmup):
        step_fn()
    per_rank_regions = []
    repeats = None
    while repeats is None or len(per_rank_regions) < repeats:
        rk.barrier()
        if device_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step_fn()
            e1.record()
            rk.barrier()
            ms = e0.elapsed_time(e1)
        else:
            t0 = time.perf_counter()
            for _ in range(steps):
                step_fn()
            torch.cuda.synchronize()
            ms = 1000.0 * (time.perf_counter() - t0)
            rk.barrier()
        per_rank_regions.append(rk.gather([ms])[:, 0])          # [world]
        if repeats is None:
            first = float(per_rank_regions[0].max())              # the same number on every rank
            repeats = int(min(max_repeats, max(1, np.ceil(min_ms / max(first, 1e-3)))))
    arr = np.stack(per_rank_regions)                              # [repeats, world]
    region_max = arr.max(1)
    med = int(np.argsort(region_max)[len(region_max) // 2])
    return float(region_max[med]), [float(v) for v in region_max], [float(v) for v in arr[med]]


def make_inputs(eng, meta, B, ensemble, rank, dev):
    """S1 candidates, seed = 1234 + rank; a 4096-candidate pool tiled to B.  Host copies are pinned."""
    from neusomatic_b200 import anns_to_255, synth
    pool = synth.uniform_pileups(4096, seed=1234 + rank, ensemble=ensemble)
    reps = (B + 4095) // 4096
    mat_h = torch.from_numpy(np.tile(pool.matrices, (reps, 1, 1, 1))[:B]).pin_memory()
    meta_h = torch.from_numpy(np.tile(pool.meta, (reps, 1))[:B].copy()).pin_memory()
    ann_h = torch.from_numpy(np.tile(anns_to_255(pool.anns), (reps, 1))[:B].copy()).pin_memory() if ensemble else None
    mat_d, meta_d = mat_h.to(dev), meta_h.to(dev)
    ann_d = ann_h.to(dev) if ensemble else None
    x_d = eng.assemble(mat_d, meta_d, ann_d, float(meta["coverage_thr"]))        # fp32 [B,C,5,32] in HBM
    torch.cuda.synchronize()
    return mat_h, meta_h, ann_h, x_d


def measure_scoring(args, rk, local_rank, ensemble, B, steps, with_f32_host=True, with_profile=True, with_clocks=True,
                    precision="arg"):
    """value (resident), e2e (host u8), e2e_f32, per-kernel profile for one model on this rank's GPU."""
    from neusomatic_b200 import Engine
    from neusomatic_b200.call import load_checkpoint
    C = 119 if ensemble else 26
    sd, meta = load_checkpoint(CKPT_ENS if ensemble else CKPT)
    eng = Engine(sd, C, device=local_rank, precision=args.precision if precision == "arg" else precision)
    mat_h, meta_h, ann_h, x_d = make_inputs(eng, meta, B, ensemble, rk.rank, rk.dev)
    thr = float(meta["coverage_thr"])
    out = {"C": C, "precision": eng.precision, "kernel": eng.kernel_name}

    # ---- value: device-resident inputs; the library's per-kernel event brackets stay on during the timed regions (two
    # event records per launch on the launching stream), so the roofline's kernel time is measured live, in the timed region
    for _ in range(args.warmup):
        eng.forward_f32(x_d, extras=True)
    sampler = ClockSampler(local_rank) if (with_clocks and rk.rank == 0) else None
    if with_profile:
        rk.barrier()
        eng.get_profile()
        eng.set_profile(True)
    l0 = eng.launch_count
    if sampler:
        sampler.start()
    ms, regions, per_rank = timed_regions(rk, lambda: eng.forward_f32(x_d, extras=True), steps, 0, min_ms=args.min_timed_ms)
    out["clocks"] = sampler.stop() if sampler else None
    out.update(ms=ms, region_ms=regions, ms_per_rank=per_rank, value=rk.world * B * steps / (ms / 1000.0))
    out["launches"] = (eng.launch_count - l0) // len(regions)          # launches of ONE K-step region
    if with_profile:
        prof = eng.get_profile()
        eng.set_profile(False)
        out["prof"] = {k: (v[0] / len(regions), v[1] // len(regions)) for k, v in prof.items()}     # per region

    # ---- e2e: host buffers through the C ABI (the call synchronises: results are in host memory when it returns) -----
    if not args.no_e2e:
        host_out = eng._host_out(B, True, pin=True)
        k = max(2, min(steps, 6))
        ms_u8, _, _ = timed_regions(rk, lambda: eng.score_host_u8(mat_h, meta_h, ann_h, thr, out=host_out), k, 1,
                                    device_events=False, min_ms=1000.0)
        d2h = B * 19 * 4
        out["e2e"] = {"value": rk.world * B * k / (ms_u8 / 1000.0), "unit": "candidates/s",
                      "h2d_bytes_per_step": B * (3680 + 12 + (372 if ensemble else 0)), "d2h_bytes_per_step": d2h,
                      "api": "nsc_score_host_u8 (pinned uint8 [B,5,32,23] + int32 [B,3] in, host results out)"}
        if with_f32_host:
            x_h = torch.empty((B, C, 5, 32), dtype=torch.float32).pin_memory()
            x_h.copy_(x_d)
            ms_f32, _, _ = timed_regions(rk, lambda: eng.score_host_f32(x_h, out=host_out), k, 1, device_events=False,
                                         min_ms=1000.0)
            out["e2e_f32"] = {"value": rk.world * B * k / (ms_f32 / 1000.0), "unit": "candidates/s",
                              "h2d_bytes_per_step": B * C * 160 * 4, "d2h_bytes_per_step": d2h,
                              "api": "nsc_score_host_f32 (pinned fp32 [B,C,5,32] in, host results out)"}
            del x_h
    eng.close()
    del x_d
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: train.py step
# ---------------------------------------------------------------------------------------------------------------------
def train_inputs(B, rank, dev, local_rank):
    """Synthetic labelled batch: S2 pileups, channel assembly by the library on the device."""
    from neusomatic_b200 import Engine, synth
    cand = synth.structured_pileups(B, seed=100 + rank)
    eng = Engine(synth.random_state_dict(26, seed=1), 26, device=local_rank)
    x = eng.assemble(torch.from_numpy(cand.matrices).to(dev), torch.from_numpy(cand.meta).to(dev), None, 100.0).clone()
    eng.close()
    vt = torch.tensor([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], dtype=torch.int32, device=dev)
    ln = torch.tensor([min(int(t.split(".")[6]), 3) for t in cand.tags], dtype=torch.int32, device=dev)
    ce = torch.from_numpy(cand.center).float().to(dev)
    w1 = torch.tensor([0.2, 0.21, 0.13, 0.16], device=dev)
    w3 = torch.tensor([0.12, 0.2, 0.23, 0.24], device=dev)
    return x, vt, ce, ln, w1, w3


def measure_train(args, rk, local_rank, B, steps):
    import torch.distributed as dist
    from neusomatic_b200 import synth
    from neusomatic_b200.train import Trainer
    x, vt, ce, ln, w1, w3 = train_inputs(B, rk.rank, rk.dev, local_rank)
    tr = Trainer(synth.random_state_dict(26, seed=11), 26, device=local_rank, max_batch=B)
    ms, regions, per_rank = timed_regions(rk, lambda: tr.step(x, vt, ce, ln, w1, w3), steps, max(args.warmup, 3), min_ms=1000.0)
    out = {"metric": "NeuSomaticNet train samples/sec (fwd+bwd+SGD)", "value": rk.world * B * steps / (ms / 1000.0), "unit": "samples/s",
           "ms_per_step": ms / steps, "region_ms": regions, "ms_per_rank": per_rank, "batch_per_gpu": B, "loss": float(tr.losses[0]),
           "allreduce_us": None,
           "dtype": "split fp16 x3 on tcgen05 with fp32 accumulate for the convolutions (forward, data and weight gradient); f32 elsewhere",
           "workload": "train.py step (BASELINE.json configs[4]): C=26, SGD lr 0.01 momentum 0.9, BatchNorm batch statistics per GPU, "
                       "one NCCL all-reduce of %d floats per step" % tr.n}
    if rk.world > 1:
        ms_ar, _, _ = timed_regions(rk, lambda: dist.all_reduce(tr.grads), 20, 3, min_ms=200.0)
        out["allreduce_us"] = 1000.0 * ms_ar / 20
    tr.close()
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# GPU incumbent: the reference architecture as a stock nn.Module in eager PyTorch on the same GPU (network.py:17-78
# layer list; cuDNN / cuBLAS kernels).  Informational: what `net.cuda()` of call.py:411-413 / train.py delivers.
# ---------------------------------------------------------------------------------------------------------------------
def stock_module(C):
    import torch.nn as nn
    import torch.nn.functional as F

    class Block(nn.Module):
        def __init__(self, k1, k2, d1, d2, st):
            super().__init__()
            self.conv_r1 = nn.Conv2d(64, 64, k1, dilation=d1, padding=(d1 * (k1 - 1)) // 2)
            self.bn_r1 = nn.BatchNorm2d(64)
            self.conv_r2 = nn.Conv2d(64, 64, k2, dilation=d2, padding=(d2 * (k2 - 1)) // 2)
            self.bn_r2 = nn.BatchNorm2d(64)
            self.pool_r2 = nn.MaxPool2d((1, 3), padding=(0, 1), stride=(1, st))

        def forward(self, x):
            y = F.relu(self.bn_r1(self.conv_r1(x)))
            return self.pool_r2(x + self.bn_r2(self.conv_r2(y)))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(C, 64, (1, 3), padding=(0, 1))
            self.bn1 = nn.BatchNorm2d(64)
            self.pool1 = nn.MaxPool2d((1, 3), padding=(0, 1), stride=(1, 1))
            self.res_layers = nn.Sequential(Block(3, 5, 1, 1, 1), Block(3, 5, 1, 1, 2), Block(3, 5, 2, 1, 2), Block(3, 5, 4, 2, 2))
            self.fc1, self.fc2, self.fc3, self.fc4 = nn.Linear(1280, 240), nn.Linear(240, 4), nn.Linear(240, 1), nn.Linear(240, 4)

        def forward(self, x):
            x = self.res_layers(self.pool1(F.relu(self.bn1(self.conv1(x)))))
            x = F.relu(self.fc1(x.reshape(x.shape[0], -1)))
            return self.fc2(x), self.fc3(x), self.fc4(x)
    return Net()


def measure_incumbent(args, rk, local_rank):
    import torch.nn.functional as F
    from neusomatic_b200 import Engine
    from neusomatic_b200.call import load_checkpoint
    dev = rk.dev
    sd, meta = load_checkpoint(CKPT)
    net = stock_module(26)
    net.load_state_dict(sd, strict=False)
    net = net.to(dev).eval()
    B = 8192
    eng = Engine(sd, 26, device=local_rank, precision=args.precision)
    _, _, _, x_d = make_inputs(eng, meta, B, False, 0, dev)
    ours = [t.float() for t in eng.forward_f32(x_d, extras=True)]
    eng.close()
    out = {"batch": B, "what": "stock nn.Module of the same architecture, eager PyTorch %s on the same GPU (cuDNN / cuBLAS), "
                               "inputs resident" % torch.__version__}
    for name, tf32 in (("fp32", False), ("tf32_allowed", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            ms, _, _ = timed_regions(rk, lambda: net(x_d), 8, 3, min_ms=500.0)
            o1, o2, o3 = net(x_d)
        p1, p3 = torch.softmax(o1, 1), torch.softmax(o3, 1)
        out["inference_" + name] = {
            "value": B * 8 / (ms / 1000.0), "unit": "candidates/s",
            "vs_this_library": {"max_dprob": float(max((p1 - ours[3]).abs().max(), (p3 - ours[4]).abs().max())),
                                "max_dpos": float((o2 - ours[1]).abs().max()),
                                "argmax_flips": int((o1.argmax(1) != ours[5]).sum() + (o3.argmax(1) != ours[6]).sum())}}
    del x_d
    # training step: loss.backward() + optim.SGD, batch 1024 (train.py:416-441)
    x, vt, ce, ln, w1, w3 = train_inputs(1024, 0, dev, local_rank)
    vt, ln = vt.long(), ln.long()
    for name, tf32 in (("fp32", False), ("tf32_allowed", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        tnet = stock_module(26).to(dev).train()
        opt = torch.optim.SGD(tnet.parameters(), lr=0.01, momentum=0.9)

        def step():
            opt.zero_grad(set_to_none=True)
            o1, o2, o3 = tnet(x)
            loss = F.cross_entropy(o1, vt, weight=w1) + F.smooth_l1_loss(o2[:, 0], ce) + F.cross_entropy(o3, ln, weight=w3)
            loss.backward()
            opt.step()
        ms, _, _ = timed_regions(rk, step, 10, 3, min_ms=500.0)
        out["train_" + name] = {"value": 1024 * 10 / (ms / 1000.0), "unit": "samples/s", "ms_per_step": ms / 10, "batch": 1024}
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------------
# e2e_tsv: reference-format candidate TSV -> dictionaries
# ---------------------------------------------------------------------------------------------------------------------
def measure_tsv(args, rk, local_rank, n=524288):
    import base64
    import zlib
    from neusomatic_b200 import synth
    from neusomatic_b200.call import CandidateScorer, PredTable
    from neusomatic_b200.vcf import pred_vcf_records_none
    pool = synth.structured_pileups(4096, seed=4321)
    td = tempfile.mkdtemp(prefix="nsc_bench_")
    path = os.path.join(td, "candidates_0.tsv")
    blobs = [base64.b64encode(zlib.compress(np.ascontiguousarray(pool.matrices[i]))).decode("ascii") for i in range(4096)]
    with open(path, "w") as f:                      # generate_dataset.py:1569-1581 line format; tags made unique by position
        for i in range(n):
            tf = pool.tags[i % 4096].split(".")
            tf[1] = str(1000 + i)
            f.write("%d\t1\t%s\t%s\n" % (i, ".".join(tf), blobs[i % 4096]))
    size = os.path.getsize(path)
    sc = CandidateScorer(CKPT, device=local_rank, precision=args.precision)
    threads = os.cpu_count() or 1
    chroms = ["chr%d" % i for i in range(30)]

    def run_mode(host_decode):
        os.environ["NSC_TSV_HOST_DECODE"] = "1" if host_decode else "0"
        sc.call_tsv(path, batch=32768, num_threads=threads)       # warm-up: page cache, workspace, pinned slots
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            fp, npred = sc.call_tsv(path, batch=32768, num_threads=threads)
            nrec = len(dict(pred_vcf_records_none(npred, chroms)))
            ts.append(time.perf_counter() - t0)
        from neusomatic_b200.call import PIPELINE_TIMES
        return float(np.median(ts)), fp, npred, nrec, {k: round(v, 4) for k, v in PIPELINE_TIMES.items()}

    el_host, _, _, _, stages_host = run_mode(True)
    el, final_preds, none_preds, n_none_records, stages = run_mode(False)
    # device time of the inflate kernel itself (event brackets of the library), one more pass
    sc.engine.set_profile(True)
    sc.engine.get_profile()
    sc.call_tsv(path, batch=32768, num_threads=threads)
    prof = sc.engine.get_profile()
    sc.engine.set_profile(False)
    os.environ.pop("NSC_TSV_HOST_DECODE", None)
    # the C-ABI call of that path alone: one staged slice of 65 536 candidates (pinned text + spans + tag fields) through
    # nsc_score_host_b64, H2D of the text, device inflate, the network and D2H of the results inside the timed region
    from neusomatic_b200.tsv import CandidateTSV
    from neusomatic_b200.engine import Engine
    t = CandidateTSV(path, num_threads=threads)
    nb = 65536
    text = torch.empty(t.range_bytes(0, nb) + 64, dtype=torch.uint8).pin_memory()
    spans = torch.empty((nb, 2), dtype=torch.int32).pin_memory()
    meta = torch.empty((nb, 3), dtype=torch.int32).pin_memory()
    t.stage(0, nb, text, spans, meta)
    nbytes = t.range_bytes(0, nb)
    t.close()
    outs = Engine._host_out(nb, True, True)
    for _ in range(3):
        sc.engine.score_host_b64(text[:nbytes], spans, meta, None, sc.coverage_thr, out=outs)
    t0 = time.perf_counter()
    for _ in range(16):
        sc.engine.score_host_b64(text[:nbytes], spans, meta, None, sc.coverage_thr, out=outs)
    b64_rate = 16 * nb / (time.perf_counter() - t0)
    sc.engine.close()
    infl_ms = prof.get("inflate_b64", (0.0, 0))[0]
    cpu_rate, cpu_sample = (None, None)
    if not args.no_cpu_baseline:
        cpu_rate, cpu_sample = cpu_tsv_rate(path, 26)
    os.unlink(path)
    os.rmdir(td)
    return {"value": n / el, "unit": "candidates/s", "candidates": n, "tsv_bytes": size, "host_threads": threads,
            "final_preds": len(final_preds), "none_preds": len(none_preds), "none_vcf_records": n_none_records,
            "what": "CandidateScorer.call_tsv: C++ line splitter / tag parser -> nsc_score_host_b64 (base64 + zlib inflate on the device, "
                    "then the network) -> PredTable dictionaries + the none.vcf filter on arrays; stage, score and filing overlapped "
                    "(call.py:504-534 + 336-354)",
            "h2d_bytes_per_candidate": size / n,
            "b64_call": {"value": b64_rate, "unit": "candidates/s", "h2d_bytes_per_step": nbytes + nb * 20, "d2h_bytes_per_step": nb * 76,
                         "api": "nsc_score_host_b64 (pinned TSV text + spans + tag fields in, host results out; inflate on the device)"},
            "seconds": el, "stage_busy_seconds": stages,
            "inflate_kernel": {"ms_per_pass": infl_ms, "candidates_per_s": (n / (infl_ms * 1e-3)) if infl_ms > 0 else None,
                               "output_GBps": (n * 3680 / (infl_ms * 1e-3) / 1e9) if infl_ms > 0 else None},
            "host_inflate": {"value": n / el_host, "unit": "candidates/s", "seconds": el_host, "stage_busy_seconds": stages_host,
                             "what": "the same call with NSC_TSV_HOST_DECODE=1: inflate on the host threads -> nsc_score_host_u8"},
            "cpu_reference_chain": None if cpu_rate is None else {
                "value": cpu_rate, "unit": "candidates/s", "kind": "port", "cores": threads,
                "sample": "oracle: TSV line decode + channel assembly + torch_port forward + call_variants loop, " + cpu_sample}}


def measure_tsv_ensemble(args, rk, local_rank, n=262144):
    """The same real-file call for the ensemble network (BASELINE configs[3]): 93 annotation fields per line, parsed by the
    staging threads; device inflate only."""
    import base64
    import zlib
    from neusomatic_b200 import synth
    from neusomatic_b200.call import CandidateScorer, PIPELINE_TIMES
    from neusomatic_b200.vcf import pred_vcf_records_none
    pool = synth.structured_pileups(4096, seed=4321, ensemble=True)
    td = tempfile.mkdtemp(prefix="nsc_bench_")
    path = os.path.join(td, "candidates_0.tsv")
    blobs = [base64.b64encode(zlib.compress(np.ascontiguousarray(pool.matrices[i]))).decode("ascii") for i in range(4096)]
    anns = ["\t".join(("%.4f" % a).rstrip("0").rstrip(".") or "0" for a in pool.anns[i]) for i in range(4096)]      # short decimals, as in real files
    with open(path, "w") as f:
        for i in range(n):
            tf = pool.tags[i % 4096].split(".")
            tf[1] = str(1000 + i)
            f.write("%d\t1\t%s\t%s\t%s\n" % (i, ".".join(tf), blobs[i % 4096], anns[i % 4096]))
    size = os.path.getsize(path)
    sc = CandidateScorer(CKPT_ENS, device=local_rank, precision=args.precision)
    threads = os.cpu_count() or 1
    chroms = ["chr%d" % i for i in range(30)]
    sc.call_tsv(path, batch=32768, num_threads=threads)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        fp, npred = sc.call_tsv(path, batch=32768, num_threads=threads)
        len(dict(pred_vcf_records_none(npred, chroms)))
        ts.append(time.perf_counter() - t0)
    sc.engine.close()
    os.unlink(path)
    os.rmdir(td)
    el = float(np.median(ts))
    return {"value": n / el, "unit": "candidates/s", "candidates": n, "tsv_bytes": size, "host_threads": threads, "seconds": el,
            "h2d_bytes_per_candidate": size / n, "stage_busy_seconds": {k: round(v, 4) for k, v in PIPELINE_TIMES.items()},
            "what": "e2e_tsv for the ensemble network: C=119, 93 annotation fields per line (BASELINE.json configs[3])"}


def measure_whole_call(args, rk, local_rank, n_none=500000, n_gen=12000):
    """The reference's program end to end (call.py:389-554): candidate TSVs + FASTA + checkpoint -> pred.vcf + none.vcf through
    ``call_neusomatic``, on a generated workload with a realistic share of called variants (tools/whole_call_probe.py)."""
    import base64
    import shutil
    import zlib
    from neusomatic_b200 import synth
    from neusomatic_b200.call import call_neusomatic
    td = tempfile.mkdtemp(prefix="nsc_bench_call_")
    try:
        cand, seqs, chroms = synth.genome_candidates(n_gen, seed=41)
        fa = os.path.join(td, "ref.fa")
        with open(fa, "w") as f:
            for c in chroms:
                f.write(">%s\n" % c)
                f.writelines(seqs[c][i:i + 60] + "\n" for i in range(0, len(seqs[c]), 60))
        pool = synth.structured_pileups(12000, seed=4321)
        sel = [i for i, t in enumerate(pool.tags) if t.split(".")[4] == "NONE"][:4096]          # pileups without a variant signal
        blobs = [base64.b64encode(zlib.compress(np.ascontiguousarray(pool.matrices[i]))).decode("ascii") for i in sel]
        tags = [pool.tags[i] for i in sel]
        gblobs = [base64.b64encode(zlib.compress(np.ascontiguousarray(cand.matrices[i]))).decode("ascii") for i in range(n_gen)]
        files, per = [], (n_none + n_gen + 3) // 4
        every = max(1, (n_none + n_gen) // max(n_gen, 1))
        k = g = 0
        for fi in range(4):
            files.append(os.path.join(td, "candidates_%d.tsv" % fi))
            with open(files[-1], "w") as f:
                for i in range(per):
                    j = fi * per + i
                    if j >= n_none + n_gen:
                        break
                    if j % every == 0 and g < n_gen:
                        f.write("%d\t1\t%s\t%s\n" % (i, cand.tags[g], gblobs[g]))
                        g += 1
                    else:
                        tf = tags[k % len(sel)].split(".")
                        tf[0], tf[1] = "0", str(10 ** 7 + k)
                        f.write("%d\t1\t%s\t%s\n" % (i, ".".join(tf), blobs[k % len(sel)]))
                        k += 1
        total = k + g
        threads = os.cpu_count() or 1
        ck = os.path.join(ROOT, "tests", "golden", "std_v014.pth")
        times = []
        for it in range(3):                              # the first call pays the workspace / pinned-slot / pool start-up
            out = os.path.join(td, "out%d" % it)
            t0 = time.perf_counter()
            call_neusomatic(files, fa, out, ck, threads, 1000, 400000, 0.7, 0.4, False, True)
            times.append(time.perf_counter() - t0)
        n_pred = sum(1 for ln in open(out + "/pred.vcf") if not ln.startswith("#"))
        n_none_rec = sum(1 for ln in open(out + "/none.vcf") if not ln.startswith("#"))
    finally:
        shutil.rmtree(td, ignore_errors=True)
    el = float(min(times[1:]))
    return {"value": total / el, "unit": "candidates/s", "candidates": total, "seconds": el, "first_call_seconds": times[0],
            "pred_vcf_records": n_pred, "none_vcf_records": n_none_rec, "host_threads": threads,
            "what": "call_neusomatic (call.py:389-554): 4 candidate TSVs + FASTA + v0.1.4 checkpoint -> pred.vcf + none.vcf; device "
                    "inflate + scoring, the VCF records of the called variants built by a process pool meanwhile"}


def cublas_bf16_same_box(dev, seconds=1.5, n=8192):
    """Informational: the rate a dense bf16 cuBLAS GEMM (n^3, torch.matmul) sustains on THIS box in THIS run -- the same
    measurement MEASURED_PEAKS.json's sustained figure comes from, but under this box's own power cap and clocks (boxes of
    the pool differ by 5-10 %).  Not on the product path; `roofline.peak` stays the driver-written number."""
    a = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    b = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    c = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
    for _ in range(20):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters, total_ms, done = 200, 0.0, 0
    while total_ms < seconds * 1000.0 and done < 20000:
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
        done += iters
    return {"tflops": 2.0 * n ** 3 * done / (total_ms / 1000.0) / 1e12, "seconds": total_ms / 1000.0, "n": n,
            "what": "torch.matmul bf16 %d^3 (cuBLAS) run back to back for >= %.1f s right after the timed regions, same process" % (n, seconds)}


def roofline_record(m, B, steps, world, C):
    """Tensor roofline of the dominant kernel family (the eight 64->64 convolutions) from the live per-kernel event times."""
    pk = peaks()
    prof = m.get("prof") or {}
    conv_ms = sum(v[0] for k_, v in prof.items() if k_.startswith("res"))
    conv_launches = sum(v[1] for k_, v in prof.items() if k_.startswith("res"))
    total_ms = sum(v[0] for v in prof.values())
    if conv_ms <= 0:
        return None
    traffic, traffic_src = None, None
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(tpath):
            t = json.load(open(tpath))
            key = {0: "dram_bytes_per_launch_split16", 3: "dram_bytes_per_launch_split8"}.get(m["precision"])
            traffic = t.get(key, t.get("dram_bytes_per_launch"))     # one ncu --set full capture, 32768-candidate launches
            traffic_src = "profiles/" + fn
            break
    cands = B * steps
    achieved = cands * 2 * CONV_STACK_MACS / (conv_ms / 1000.0) / 1e12
    slots = MMA_SLOTS.get(m["precision"], 1)
    executed = cands * 2 * CONV_STACK_EXEC_MACS * slots / (conv_ms / 1000.0) / 1e12 if m["precision"] != 2 else None
    return {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_sustained"], "peak_burst": pk["bf16_burst"], "frac_burst": achieved / pk["bf16_burst"],
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["src"] + " bf16 (sustained = frac, burst = frac_burst)",
            "kernel": m["kernel"], "kernel_launches": conv_launches,
            "kernel_ms_per_launch": conv_ms / max(conv_launches, 1),
            "kernel_share_of_step": conv_ms / total_ms if total_ms else None,
            "algorithmic_flops_per_candidate": 2 * CONV_STACK_MACS,
            # the parity arithmetic issues `slots` tensor instructions per algorithmic MAC (3 fp16 products, or 1 fp16 + 2 fp8 at
            # twice the K) and skips the 24 % of kernel rows in the H padding: 100 % tensor-pipe occupancy = frac_ceiling
            "mma_slots_per_mac": slots, "frac_ceiling_at_full_tensor_pipe": CONV_STACK_MACS / (CONV_STACK_EXEC_MACS * slots),
            "executed_tensor_tflops_fp16_slots": executed,
            "whole_net_frac": (m["value"] / world) * FLOPS_PER_CAND[C] / 1e12 / pk["bf16_sustained"],
            "whole_net_frac_burst": (m["value"] / world) * FLOPS_PER_CAND[C] / 1e12 / pk["bf16_burst"],
            "per_kernel_ms": {k_: round(v[0] / steps, 4) for k_, v in prof.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="candidates per GPU per step")
    ap.add_argument("--precision", type=int, default=None, help="NSC_PREC_* (default: library default)")
    ap.add_argument("--ref-batch", type=int, default=1000)
    ap.add_argument("--ref-seconds", type=float, default=None, help="--impl reference: CPU seconds per step (default: 120 / steps, "
                    "clamped to 3..20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--min-timed-ms", type=float, default=MIN_TIMED_MS, help="repeat the K-step region until this much is timed "
                    "(0 = one region: profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra records (ensemble, train, gpu_incumbent, e2e_tsv)")
    ap.add_argument("--train", action="store_true", help="BASELINE configs[4] as the line's own metric")
    ap.add_argument("--ensemble", action="store_true", help="BASELINE configs[3] as the line's own metric (C=119)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL writes its version banner / debug lines to stdout by default: send them to stderr, stdout is the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    rk = Ranks(rank, world, dev)

    if args.train:
        t = measure_train(args, rk, local_rank, min(args.batch, 1024), args.steps)
        if rank == 0:
            print(json.dumps({"metric": t["metric"], "value": t["value"], "unit": t["unit"], "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": t["dtype"], "data": "synthetic",
                              "config": {"workload": t["workload"], "batch_per_gpu": t["batch_per_gpu"]}, "loss": t["loss"],
                              "allreduce_us": t["allreduce_us"], "region_ms": t["region_ms"], "ms_per_rank": t["ms_per_rank"]}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    B = args.batch
    m = measure_scoring(args, rk, local_rank, args.ensemble, B, args.steps)
    C = m["C"]
    same = None
    if not args.no_extra:
        try:                         # right after the timed regions: the GPU is in the power / thermal state they ran in
            same = cublas_bf16_same_box(dev)
        except Exception as ex:      # noqa: BLE001  (informational)
            same = {"error": "%s: %s" % (type(ex).__name__, ex)}

    extra = {}
    if not args.no_extra:
        if not args.ensemble:
            e = measure_scoring(args, rk, local_rank, True, min(B, 32768), max(4, args.steps // 2), with_f32_host=False,
                                with_profile=False, with_clocks=False)
            extra["ensemble"] = {"value": e["value"], "unit": "candidates/s", "ms_per_step": e["ms"] / max(4, args.steps // 2),
                                 "batch_per_gpu": min(B, 32768), "e2e": e.get("e2e"), "ms_per_rank": e["ms_per_rank"],
                                 "workload": "synthetic ensemble call (BASELINE.json configs[3]): C=119, v0.1.0 ensemble checkpoint"}
        if args.precision is None:      # the opt-in fast arithmetic (NSC_PREC_SPLIT8) on the same workload, same run
            f = measure_scoring(args, rk, local_rank, args.ensemble, B, args.steps, with_f32_host=False, with_clocks=False, precision=3)
            fr = roofline_record(f, B, args.steps, world, C)
            extra["fast_mode"] = {"dtype": PREC_NAMES[3], "value": f["value"], "unit": "candidates/s", "ms_per_step": f["ms"] / args.steps,
                                  "e2e": f.get("e2e"), "ms_per_rank": f["ms_per_rank"],
                                  "roofline": None if fr is None else {k: fr[k] for k in ("achieved", "frac", "frac_burst", "frac_ceiling_at_full_tensor_pipe",
                                                                                        "kernel_share_of_step", "per_kernel_ms")},
                                  "parity": "meets the north-star tolerances (argmax >= 99.99 %, 2e-3, 1e-3) with 10x margin; 1-8 % of the 4-decimal "
                                            "probabilities differ from the reference's (default mode: <= 1 %), profiles/r2_parity_large.json"}
        t = measure_train(args, rk, local_rank, 1024, 20)
        extra["train"] = {k: t[k] for k in ("value", "unit", "ms_per_step", "batch_per_gpu", "allreduce_us", "ms_per_rank", "workload")}
        if world == 1:
            # informational records: a failure in one of them (a full /tmp, say) must not cost the line its headline numbers
            for name, fn in (("gpu_incumbent", measure_incumbent), ("e2e_tsv", measure_tsv),
                             ("e2e_tsv_ensemble", measure_tsv_ensemble), ("whole_call", measure_whole_call)):
                try:
                    extra[name] = fn(args, rk, local_rank)
                except Exception as ex:      # noqa: BLE001
                    extra[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}

    roof = roofline_record(m, B, args.steps, world, C)
    if roof is not None and same is not None:
        roof["cublas_bf16_same_box"] = same
        if "tflops" in same:
            roof["frac_of_cublas_same_box"] = roof["achieved"] / same["tflops"]
            if roof.get("executed_tensor_tflops_fp16_slots"):
                roof["executed_vs_cublas_same_box"] = roof["executed_tensor_tflops_fp16_slots"] / same["tflops"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, threads, sample, _ = cpu_port_rate(C, seconds=12.0, batch=args.ref_batch)
        cpu = {"value": r, "unit": "candidates/s", "cores": threads, "kind": "port",
               "sample": "oracle/torch_port.py (PyTorch CPU operators, fp32): " + sample}

    if rank == 0:
        prec_names = PREC_NAMES
        print(json.dumps({
            "metric": "NeuSomaticNet candidates/sec", "value": m["value"], "unit": "candidates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec_names.get(m["precision"], str(m["precision"])), "data": "synthetic",
            "config": workload_config(args, world),
            "timing": "median of %d regions of %d steps (>= %.0f ms timed), max over ranks per region" % (
                len(m["region_ms"]), args.steps, args.min_timed_ms),
            "region_ms": [round(v, 3) for v in m["region_ms"]], "ms_per_rank": [round(v, 3) for v in m["ms_per_rank"]],
            "e2e": m.get("e2e"), "e2e_f32": m.get("e2e_f32"), "gpu_launches": m.get("launches"), "roofline": roof,
            "cpu_baseline": cpu, "clocks": m["clocks"], "extra": extra,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
