"""Where the time of a SMALL device call goes (BASELINE.json configs[2], C = 26, loader batches of 100 / 1000 / 2048).
For each batch size, with the batch resident in HBM:
  * issue_us  : host time of one ``Engine.forward_f32`` call (ctypes + the library's launches), GPU not waited for;
  * device_us : CUDA-event time per call of a back-to-back loop (what the sweep of tools/batch_sweep.py reports);
  * single_us : CUDA-event time of ONE call issued on an idle GPU (latency of the chain of dependent launches);
  * kernels   : the library's per-kernel event brackets (``nsc_model_set_profile``), microseconds per launch.
Run it under ``ncu --metrics gpu__time_duration.sum`` as well for the cold-cache per-launch durations."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neusomatic_b200 import Engine, synth  # noqa: E402

BATCHES = [100, 1000, 2048]


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dev = torch.device("cuda", 0)
    sd = synth.random_state_dict(26, seed=7)
    eng = Engine(sd, 26)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    res = {"iters": iters, "precision": eng.precision, "rows": []}
    for B in BATCHES:
        x = torch.rand((B, 26, 5, 32), device=dev, generator=g)
        for _ in range(5):
            eng.forward_f32(x, extras=True)
        torch.cuda.synchronize()
        # host issue time
        t0 = time.perf_counter()
        for _ in range(iters):
            eng.forward_f32(x, extras=True)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        # back-to-back device time
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            eng.forward_f32(x, extras=True)
        e1.record()
        torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) / iters * 1e3
        # one call on an idle GPU
        singles = []
        for _ in range(20):
            torch.cuda.synchronize()
            e0.record()
            eng.forward_f32(x, extras=True)
            e1.record()
            torch.cuda.synchronize()
            singles.append(e0.elapsed_time(e1) * 1e3)
        # per-kernel brackets
        eng.set_profile(True)
        eng.get_profile()
        for _ in range(50):
            eng.forward_f32(x, extras=True)
        prof = eng.get_profile()
        eng.set_profile(False)
        row = {"batch": B, "issue_us": (t1 - t0) / iters * 1e6, "device_us": dev_us, "single_us": float(np.median(singles)),
               "launches_per_call": None,
               "kernels_us": {k: round(v[0] / max(v[1], 1) * 1e3, 2) for k, v in prof.items()},
               "kernels_sum_us": round(sum(v[0] for v in prof.values()) / 50 * 1e3, 1)}
        c0 = eng.launch_count
        eng.forward_f32(x, extras=True)
        row["launches_per_call"] = eng.launch_count - c0
        res["rows"].append(row)
        print(json.dumps(row), flush=True)
    eng.close()
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "small_batch_probe.json"), "w") as f:
            f.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
