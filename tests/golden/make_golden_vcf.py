#!/usr/bin/env python
"""Golden fixtures for the dictionaries -> VCF stage, produced by the UNMODIFIED reference ``call.py``.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_vcf.py

The reference's ``call.py`` imports ``pysam`` and ``imageio`` (and ``merge_tsvs`` imports ``pybedtools``), none of which
is installed here.  They are replaced by small stand-ins registered in ``sys.modules`` BEFORE the reference is imported
-- the reference sources themselves are imported from /root/reference untouched:

  * ``pysam.FastaFile``   -- reads a plain FASTA file: ``references``, ``fetch(chrom, start, end)`` (0-based, half open),
                             context manager.  That is all call.py / utils.py use (call.py:125,276,403-405; utils.py:65-69).
  * ``imageio.imwrite / imread`` -- a real PNG round trip through Pillow (call.py:92-93,128).
  * ``pybedtools``        -- empty module (only merge_tsvs' splitting of oversized TSVs would use it; not reached).
  * numpy >= 2: ``dataloader.extract_zlib`` uses the removed ``np.fromstring``; same one-line shim as make_golden.py.

Two kinds of fixtures are written to tests/golden/vcf_<name>.npz:

  1. END TO END: the reference's own ``call_neusomatic`` (CPU) on seeded ``synth.genome_candidates`` written as two
     reference-format candidate TSVs plus a synthetic FASTA: the resulting ``pred.vcf`` / ``none.vcf`` text, and the
     (final_preds, none_preds) dictionaries its ``call_variants`` returned (captured by wrapping the function, not by
     editing it).
  2. FORCED: ``pred_vcf_records_path`` / ``pred_vcf_records_none`` / ``write_vcf`` called directly on dictionaries made
     from the planted labels with perturbed centres and probabilities, so that the SNP / INS / DEL record builders, the
     centre refinement loop and every early return are exercised whatever the bundled weights predict on synthetic data.
"""
import json
import os
import pickle
import sys
import tempfile
import types
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/neusomatic"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "python"))


# ---- stand-ins for the missing third-party modules --------------------------------------------------------------
class _FastaFile:
    def __init__(self, path):
        self.seqs, name = {}, None
        for line in open(path):
            line = line.strip()
            if line.startswith(">"):
                name = line[1:].split()[0]
                self.seqs[name] = []
            elif line:
                self.seqs[name].append(line)
        self.seqs = {k: "".join(v) for k, v in self.seqs.items()}
        self.references = tuple(self.seqs.keys())

    def fetch(self, reference, start, end):
        return self.seqs[reference][max(start, 0):end]

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _install_stubs():
    from PIL import Image
    pysam = types.ModuleType("pysam")
    pysam.FastaFile = _FastaFile
    imageio = types.ModuleType("imageio")
    imageio.imwrite = lambda fn, arr: Image.fromarray(np.asarray(arr, np.uint8)).save(fn)
    imageio.imread = lambda fn: np.array(Image.open(fn))
    sys.modules["pysam"] = pysam
    sys.modules["imageio"] = imageio
    sys.modules["pybedtools"] = types.ModuleType("pybedtools")


_install_stubs()
import dataloader as ref_dataloader            # noqa: E402  (reference, unmodified)
import call as ref_call                        # noqa: E402  (reference, unmodified)

ref_dataloader.extract_zlib = lambda b: np.frombuffer(zlib.decompress(b), np.uint8).reshape(5, 32, 23).copy()

from neusomatic_b200 import synth              # noqa: E402
from oracle import nsnet_oracle as oracle      # noqa: E402

CASES = {
    # name: (bundled checkpoint, ensemble, candidates, seed)
    # (the same bundled checkpoints as tests/golden/<name>.pth; the v0.1.0 Dream3 model calls NONE on nearly every
    # synthetic pileup, so the cases use the models that do call variants on them)
    "std_v014": ("NeuSomatic_v0.1.4_standalone_SEQC-WGS-Spike.pth", False, 360, 3101),
    "std_wex": ("NeuSomatic_v0.1.0_standalone_WEX_100purity.pth", False, 240, 3103),
    "ens_v014": ("NeuSomatic_v0.1.4_ensemble_SEQC-WGS-GT50-SpikeWGS10.pth", True, 180, 3102),
}


def write_fasta(path, fasta):
    with open(path, "w") as f:
        for name, seq in fasta.items():
            f.write(">%s\n" % name)
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + "\n")


def write_tsv(path, cand, lo, hi):
    offs = []
    with open(path, "w") as f:
        for i in range(lo, hi):
            offs.append(f.tell())
            f.write(oracle.encode_tsv_line(i - lo, cand.tags[i], cand.matrices[i], None if cand.anns is None else cand.anns[i]))
        offs.append(f.tell())
    pickle.dump(offs, open(path + ".idx", "wb"))


def dict_arrays(prefix, d):
    """call_variants dictionary -> flat arrays (values are np.float32 scalars rounded by the reference)."""
    keys = sorted(d.keys())
    out = {prefix + "_paths": np.array(keys)}
    if not keys:
        keys = []
    out[prefix + "_type"] = np.array([d[k][0] for k in keys])
    out[prefix + "_pos"] = np.array([np.asarray(d[k][1], np.float32).reshape(-1)[0] for k in keys], np.float32)
    out[prefix + "_len"] = np.array([int(d[k][2]) for k in keys], np.int64)
    for j, nm in ((3, "p1"), (4, "p3"), (5, "o1"), (6, "o3")):
        out[prefix + "_" + nm] = np.array([[np.float32(v) for v in d[k][j]] for k in keys], np.float32).reshape(len(keys), 4)
    return out


def forced_dicts(cand, seed):
    """Predictions made from the planted labels: type = label (sometimes another class), centre = planted centre plus
    noise of up to +-4 columns, length class = min(length, 3) (sometimes off by one), probabilities random but consistent."""
    rng = np.random.Generator(np.random.PCG64(seed))
    final_preds = {}
    classes = oracle.VARTYPE_CLASSES
    for i, tag in enumerate(cand.tags):
        f = tag.split(".")
        t = classes.index(f[4])
        if rng.random() < 0.15:
            t = int(rng.integers(0, 4))
        if classes[t] == "NONE":
            t = 3
        p1 = rng.dirichlet([0.3] * 4).astype(np.float32)
        p1[[t, int(p1.argmax())]] = p1[[int(p1.argmax()), t]]            # type t has the largest probability
        if rng.random() < 0.1:
            p1 = (p1 * np.float32(0.04)).astype(np.float32)                # sum below min_acceptable_probmax
        ln = min(int(f[6]), 3)
        if rng.random() < 0.2:
            ln = int(np.clip(ln + rng.integers(-1, 2), 0, 3))
        p3 = rng.dirichlet([0.3] * 4).astype(np.float32)
        p3[[ln, int(p3.argmax())]] = p3[[int(p3.argmax()), ln]]
        centre = np.float32(int(f[5]) + rng.choice([0.0, 0.3, -0.6, 1.2, -2.4, 3.6, -4.5, 40.0], p=[.4, .15, .15, .1, .08, .05, .05, .02]))
        o1 = np.log(np.maximum(p1, 1e-6)).astype(np.float32)
        o3 = np.log(np.maximum(p3, 1e-6)).astype(np.float32)
        final_preds[tag] = [classes[t], np.array([centre], np.float32), np.int64(ln),
                            [round(v, 4) for v in p1], [round(v, 4) for v in p3],
                            [round(v, 4) for v in o1], [round(v, 4) for v in o3]]
    return final_preds


def main():
    torch.set_num_threads(1)
    for name, (fn, ensemble, n, seed) in CASES.items():
        torch.manual_seed(seed)
        np.random.seed(seed)
        cand, fasta, chroms = synth.genome_candidates(n, seed=seed, ensemble=ensemble)
        captured = []
        orig = ref_call.call_variants

        def capture(*a, **k):
            r = orig(*a, **k)
            captured.append(r)
            return r
        ref_call.call_variants = capture
        with tempfile.TemporaryDirectory() as td:
            ref_fa = os.path.join(td, "ref.fa")
            write_fasta(ref_fa, fasta)
            tsvs = [os.path.join(td, "candidates_0.tsv"), os.path.join(td, "candidates_1.tsv")]
            write_tsv(tsvs[0], cand, 0, n // 3)
            write_tsv(tsvs[1], cand, n // 3, n)
            out_dir = os.path.join(td, "out")
            # max_load_candidates = 2000: every file is below max/2 (no splitting) and above max/10 (one group each,
            # call.py:504-507), so the loop of call.py:500-534 runs twice
            vcf = ref_call.call_neusomatic(tsvs, ref_fa, out_dir, os.path.join(REF, "models", fn), 1, 100, 2000, 0.7, 0.4,
                                           ensemble, False)
            pred_vcf = open(vcf).read()
            none_vcf = open(os.path.join(out_dir, "none.vcf")).read()
            ref_call.call_variants = orig
            final_preds, none_preds = {}, {}
            for f_, n_, _tp in captured:
                final_preds.update(f_)
                none_preds.update(n_)
            # ---- forced dictionaries through the reference's record builders
            forced = forced_dicts(cand, seed + 1)
            nt = oracle.non_transformed(cand.matrices, cand.center)
            mat_dir = os.path.join(td, "forced_png")
            os.mkdir(mat_dir)
            recs = []
            classes = oracle.VARTYPE_CLASSES
            for i, tag in enumerate(cand.tags):
                png = os.path.join(mat_dir, "%d.png" % i)
                sys.modules["imageio"].imwrite(png, nt[i])
                r = ref_call.pred_vcf_records_path([tag, png, forced[tag], chroms, classes, ref_fa])
                assert r is not None, "reference pred_vcf_records_path raised for %s" % tag
                recs.append(r)
            forced_records = dict(filter(None, recs))
            forced_vcf_path = os.path.join(td, "forced.vcf")
            ref_call.write_vcf(ref_call.get_vcf_records(forced_records), forced_vcf_path,
                               {c: i for i, c in enumerate(chroms)}, 0.7, 0.4)
            forced_vcf = open(forced_vcf_path).read()
            forced_none = dict(ref_call.pred_vcf_records_none(forced, chroms))
            forced_none_path = os.path.join(td, "forced_none.vcf")
            ref_call.write_vcf(ref_call.get_vcf_records(forced_none), forced_none_path,
                               {c: i for i, c in enumerate(chroms)}, 0.7, 0.4)
            forced_none_vcf = open(forced_none_path).read()
        out = dict(matrices=cand.matrices, center=cand.center, tumor_cov=cand.tumor_cov, normal_cov=cand.normal_cov,
                   tags=np.array(cand.tags), chroms=np.array(chroms),
                   fasta=np.array(json.dumps(fasta)), pred_vcf=np.array(pred_vcf), none_vcf=np.array(none_vcf),
                   forced_vcf=np.array(forced_vcf), forced_none_vcf=np.array(forced_none_vcf), split=np.int64(n // 3))
        if cand.anns is not None:
            out["anns"] = cand.anns
        out.update(dict_arrays("final", final_preds))
        out.update(dict_arrays("none", none_preds))
        out.update(dict_arrays("forced", forced))
        np.savez_compressed(os.path.join(HERE, "vcf_" + name + ".npz"), **out)
        print(name, "candidates", n, "final", len(final_preds), "none", len(none_preds),
              "pred.vcf lines", pred_vcf.count("\n"), "none.vcf lines", none_vcf.count("\n"),
              "| forced records", len(forced_records), "forced.vcf lines", forced_vcf.count("\n"),
              "forced none lines", forced_none_vcf.count("\n"))
        print(pred_vcf[:400])


if __name__ == "__main__":
    main()
