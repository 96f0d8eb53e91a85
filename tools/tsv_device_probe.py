"""One pass of CandidateScorer.call_tsv over a generated reference-format TSV (device inflate), for profilers:
``ncu --metrics gpu__time_duration.sum -k regex:inflate_b64 python tools/tsv_device_probe.py 65536`` lists the decoder
kernel's launches alone (serialised, so without the convolution kernels it shares the GPU with in a real pass)."""
import base64
import os
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neusomatic_b200 import synth  # noqa: E402
from neusomatic_b200.call import CandidateScorer, PIPELINE_TIMES, PredTable, run_tsv_pipeline  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    batches = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [32768]
    pool = synth.structured_pileups(4096, seed=4321)
    td = tempfile.mkdtemp(prefix="nsc_probe_")
    path = os.path.join(td, "candidates_0.tsv")
    blobs = [base64.b64encode(zlib.compress(np.ascontiguousarray(pool.matrices[i]))).decode("ascii") for i in range(4096)]
    with open(path, "w") as f:
        for i in range(n):
            tf = pool.tags[i % 4096].split(".")
            tf[1] = str(1000 + i)
            f.write("%d\t1\t%s\t%s\n" % (i, ".".join(tf), blobs[i % 4096]))
    ck = os.path.join(ROOT, "tests", "golden", "std_v010.pth")
    import torch
    n_dev = int(sys.argv[4]) if len(sys.argv) > 4 else 1          # GPUs driven by the one pipeline (slices round-robin)
    n_dev = min(n_dev, torch.cuda.device_count())
    scorers = [CandidateScorer(ck, device=d) for d in range(n_dev)]
    for batch in batches:
        for _ in range(passes):
            t0 = time.perf_counter()
            fp, npred = PredTable(), PredTable()
            run_tsv_pipeline(scorers, [(path, 0, None)], batch, os.cpu_count() or 1, fp, npred)
            dt = time.perf_counter() - t0
            print("%d GPU(s), slices of %d: %d candidates in %.3f s = %.0f candidates/s; stages %s" % (
                n_dev, batch, n, dt, n / dt, {k: round(v, 4) for k, v in PIPELINE_TIMES.items()}), flush=True)
    for sc in scorers:
        sc.engine.close()
    os.unlink(path)
    os.rmdir(td)


if __name__ == "__main__":
    main()
