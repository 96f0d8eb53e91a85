"""The WHOLE call -- candidate TSVs + FASTA + checkpoint -> pred.vcf / none.vcf through ``call_neusomatic`` (reference
call.py:389-554) -- timed on a generated workload with a realistic share of called variants:
    python tools/whole_call_probe.py [none_candidates=500000] [genome_candidates=12000] [threads=os.cpu_count()]
``genome_candidates`` pileups agree with a synthetic genome (about two thirds of them are called and pass the VCF stage's
checks); ``none_candidates`` are structured pileups without a variant signal (tag type NONE), which the model mostly calls NONE.  Prints seconds and
candidates/s for num_threads = 1 (records built in the calling process) and num_threads = threads (process pool, fed while the
next group is scored)."""
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
from neusomatic_b200.call import call_neusomatic  # noqa: E402


def main():
    n_none = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
    n_gen = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    td = tempfile.mkdtemp(prefix="nsc_call_")
    cand, seqs, chroms = synth.genome_candidates(n_gen, seed=41)
    fa = os.path.join(td, "ref.fa")
    with open(fa, "w") as f:
        for c in chroms:
            f.write(">%s\n" % c)
            f.writelines(seqs[c][i:i + 60] + "\n" for i in range(0, len(seqs[c]), 60))
    pool = synth.structured_pileups(12000, seed=4321)
    sel = [i for i, t in enumerate(pool.tags) if t.split(".")[4] == "NONE"][:4096]      # pileups without a variant signal
    pool.tags = [pool.tags[i] for i in sel]
    pool.matrices = pool.matrices[sel]
    assert len(sel) == 4096
    blobs = [base64.b64encode(zlib.compress(np.ascontiguousarray(pool.matrices[i]))).decode("ascii") for i in range(4096)]
    gblobs = [base64.b64encode(zlib.compress(np.ascontiguousarray(cand.matrices[i]))).decode("ascii") for i in range(n_gen)]
    # four files; the genome candidates are spread evenly among the NONE-like ones (chromosome index 0 exists in the FASTA)
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
                    tf = pool.tags[k % 4096].split(".")
                    tf[0], tf[1] = "0", str(10 ** 7 + k)
                    f.write("%d\t1\t%s\t%s\n" % (i, ".".join(tf), blobs[k % 4096]))
                    k += 1
    total = k + g
    import logging
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    ck = os.path.join(ROOT, "tests", "golden", "std_v014.pth")
    for nt in (threads, 1, threads):
        out = os.path.join(td, "out_%d_%d" % (nt, int(time.time() * 1000) % 100000))
        t0 = time.perf_counter()
        call_neusomatic(files, fa, out, ck, nt, 1000, 400000, 0.7, 0.4, False, True)
        dt = time.perf_counter() - t0
        print("num_threads %2d: %d candidates (%d genome-consistent) -> pred.vcf %d records, none.vcf %d records in %.3f s = %.0f candidates/s" % (
            nt, total, g, sum(1 for ln in open(out + "/pred.vcf") if not ln.startswith("#")),
            sum(1 for ln in open(out + "/none.vcf") if not ln.startswith("#")), dt, total / dt), flush=True)


if __name__ == "__main__":
    main()
