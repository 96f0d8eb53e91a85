"""BASELINE.json configs[2]: synthetic standalone call (C = 26, bundled-layout weights) on 1 B200, batch sweep
100 .. 8192 (+ the 65 536 of bench.py).  For every batch size B the drop-in module's eval forward
(``NeuSomaticNet.forward`` -> custom op -> ``nsc_forward_f32``; call.py:76-77 is this call) scores the same
candidate pool batch by batch, the way call_variants iterates its DataLoader:
  * resident : the pool is in HBM (fp32 [N,26,5,32]); CUDA events around the loop;
  * host_u8  : ``CandidateScorer``/``nsc_score_host_u8`` on pinned uint8 batches, H2D and D2H inside the timed region.
Prints one JSON object (also written to gpurun_out/batch_sweep.json when that directory exists)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neusomatic_b200 import Engine, synth  # noqa: E402
from neusomatic_b200.network import NeuSomaticNet  # noqa: E402
from neusomatic_b200.call import call_variants  # noqa: E402
from neusomatic_b200.engine import VARTYPE_CLASSES  # noqa: E402

BATCHES = [100, 256, 512, 1000, 2048, 4096, 8192, 65536]


def main():
    pool = int(sys.argv[1]) if len(sys.argv) > 1 else 262144        # candidates per measurement (4.4 GB fp32 > L2)
    dev = torch.device("cuda", 0)
    sd = synth.random_state_dict(26, seed=7)
    net = NeuSomaticNet(26)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    net = net.to(dev).eval()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    x = torch.rand((pool, 26, 5, 32), device=dev, generator=g)
    cand = synth.uniform_pileups(65536, seed=1234) if hasattr(synth, "uniform_pileups") else synth.structured_pileups(65536, seed=1234)
    mats = torch.from_numpy(cand.matrices).pin_memory()
    meta = torch.from_numpy(np.stack([cand.center, cand.tumor_cov, cand.normal_cov], 1).astype(np.int32)).pin_memory()
    eng = Engine(sd, 26)
    res = {"pool": pool, "rows": []}
    with torch.no_grad():
        for B in BATCHES:
            nb = max(pool // B, 1)
            for _ in range(3):
                net(x[:B])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(nb):
                net(x[i * B:(i + 1) * B])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            # end to end from pinned host uint8 (65 536-candidate pool, batch by batch)
            nb2 = max(65536 // B, 1)
            host_out = eng._host_out(B, True, pin=True)          # pinned result buffers, allocated once per batch size
            eng.score_host_u8(mats[:B], meta[:B], None, 100.0, out=host_out)
            t0 = time.perf_counter()
            for i in range(nb2):
                eng.score_host_u8(mats[i * B:(i + 1) * B], meta[i * B:(i + 1) * B], None, 100.0, out=host_out)
            t1 = time.perf_counter()
            # the drop-in call_variants loop (call.py:54-118) over loader-shaped batches of B pinned fp32 matrices: H2D per
            # batch, scoring in coalesced slices, dictionaries built on the host
            nb3 = max(min(65536, 16 * 8192) // B, 1)
            xh = x[:min(nb3 * B, pool)].cpu().pin_memory()
            nt = torch.zeros((B, 5, 32, 3), dtype=torch.uint8)
            tags = ["1.%d.A.C.SNP.16.1.50.50" % i for i in range(xh.shape[0])]
            loader = [((xh[i * B:(i + 1) * B], None, None, None, nt), (tags[i * B:(i + 1) * B], None)) for i in range(xh.shape[0] // B)]
            call_variants(net, VARTYPE_CLASSES, loader[:(2 * 8192) // B + 2], None, "t", True)     # warm-up incl. one full coalesced slice
            t2 = time.perf_counter()
            call_variants(net, VARTYPE_CLASSES, loader, None, "t", True)
            t3 = time.perf_counter()
            row = {"batch": B, "resident_cand_per_s": nb * B / ms * 1e3, "resident_us_per_batch": ms / nb * 1e3,
                   "host_u8_cand_per_s": nb2 * B / (t1 - t0), "host_u8_us_per_batch": (t1 - t0) / nb2 * 1e6,
                   "call_variants_cand_per_s": len(loader) * B / (t3 - t2)}
            res["rows"].append(row)
            print("B=%6d  resident %10.0f cand/s (%8.1f us/batch)   host u8 %10.0f cand/s (%8.1f us/batch)   call_variants loop %10.0f cand/s" % (
                B, row["resident_cand_per_s"], row["resident_us_per_batch"], row["host_u8_cand_per_s"], row["host_u8_us_per_batch"],
                row["call_variants_cand_per_s"]), flush=True)
    eng.close()
    out = json.dumps(res)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "batch_sweep.json"), "w") as f:
            f.write(out + "\n")
    print(out)


if __name__ == "__main__":
    main()
