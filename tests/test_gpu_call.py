"""GPU: the whole call path -- candidate TSVs + FASTA + checkpoint -> pred.vcf / none.vcf through call_neusomatic
(call.py:389-554) -- against the files the UNMODIFIED reference wrote for the same inputs on CPU
(tests/golden/vcf_*.npz, tests/golden/make_golden_vcf.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, vcf_identity
from oracle import nsnet_oracle as oracle

pytestmark = pytest.mark.gpu


def _write_inputs(tmp_path, name):
    g = dict(np.load(os.path.join(GOLDEN, "vcf_%s.npz" % name), allow_pickle=False))
    tags = [str(t) for t in g["tags"]]
    anns = g.get("anns")
    split = int(g["split"])
    tsvs = [str(tmp_path / "candidates_0.tsv"), str(tmp_path / "candidates_1.tsv")]
    for p, (lo, hi) in zip(tsvs, ((0, split), (split, len(tags)))):
        with open(p, "w") as f:
            for i in range(lo, hi):
                f.write(oracle.encode_tsv_line(i - lo, tags[i], g["matrices"][i], None if anns is None else anns[i]))
    fa = str(tmp_path / "ref.fa")
    with open(fa, "w") as f:
        for k, s in json.loads(str(g["fasta"])).items():
            f.write(">%s\n" % k)
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    return g, tsvs, fa


# (precision, minimum share of byte-identical pred.vcf lines).  FP32 mode: only last-bit differences of fp32 summation
# order are left (the reference differs from itself by as much between batch sizes, tests/test_call_host.py).  Default mode
# (split fp16 + fp8 corrections): probabilities move by up to ~1e-4, i.e. some 4-decimal SCOREs by one unit; the records
# (chrom, pos, ref, alt), their order and the FILTER column must still be the reference's (vcf_identity asserts that).
@pytest.mark.parametrize("precision,min_same", [(2, 0.97), (3, 0.75), (0, 0.93)])
@pytest.mark.parametrize("name", ["std_v014", "std_wex", "ens_v014"])
def test_call_neusomatic_writes_the_reference_vcf(tmp_path, monkeypatch, name, precision, min_same):
    from neusomatic_b200.call import call_neusomatic
    g, tsvs, fa = _write_inputs(tmp_path, name)
    monkeypatch.setenv("NSC_PRECISION", str(precision))
    out_dir = str(tmp_path / "out")
    vcf_path = call_neusomatic(tsvs, fa, out_dir, os.path.join(GOLDEN, name + ".pth"), 2, 100, 2000, 0.7, 0.4,
                               "anns" in g, True)
    assert vcf_path == out_dir + "/pred.vcf"
    same, total = vcf_identity(open(vcf_path).read(), str(g["pred_vcf"]))
    same_n, total_n = vcf_identity(open(out_dir + "/none.vcf").read(), str(g["none_vcf"]))
    print("%s precision %d: pred.vcf %d/%d lines identical, none.vcf %d/%d" % (name, precision, same, total, same_n, total_n))
    assert total == str(g["pred_vcf"]).count("\n")
    if total >= 50:          # (the ensemble case calls 2 variants on these synthetic pileups: a share of 2 lines says nothing;
        assert same >= min_same * total      #  vcf_identity has already checked records, order and SCORE tolerance)


def test_call_neusomatic_needs_cuda_flag(tmp_path):
    from neusomatic_b200.call import call_neusomatic
    with pytest.raises(RuntimeError):
        call_neusomatic([], "x.fa", str(tmp_path), os.path.join(GOLDEN, "std_v014.pth"), 1, 100, 1000, 0.7, 0.4, False, False)


def test_call_neusomatic_with_the_record_pool(tmp_path):
    """A call large enough for the process pool of the VCF stage (call.py:308-333): several groups, > 2048 called variants,
    the pool forked after CUDA is up and fed while the next group is scored.  Same pred.vcf / none.vcf bytes as the same call
    with the records built in the calling process."""
    from neusomatic_b200 import synth
    from neusomatic_b200.call import call_neusomatic
    n = 24000
    cand, seqs, chroms = synth.genome_candidates(n, seed=31)
    fa = str(tmp_path / "ref.fa")
    with open(fa, "w") as f:
        for c in chroms:
            f.write(">%s\n" % c)
            f.writelines(seqs[c][i:i + 60] + "\n" for i in range(0, len(seqs[c]), 60))
    tsvs = []
    for k, (lo, hi) in enumerate(((0, 9000), (9000, 17000), (17000, n))):
        tsvs.append(str(tmp_path / ("candidates_%d.tsv" % k)))
        with open(tsvs[-1], "w") as f:
            for i in range(lo, hi):
                f.write(oracle.encode_tsv_line(i - lo, cand.tags[i], cand.matrices[i]))
    ck = os.path.join(GOLDEN, "std_v014.pth")
    outs = {}
    for threads in (1, 4):
        out_dir = str(tmp_path / ("out%d" % threads))
        # max_load_candidates 16000: every file is its own piece, groups end after > 1600 candidates -> three groups
        call_neusomatic(tsvs, fa, out_dir, ck, threads, 100, 16000, 0.7, 0.4, False, True)
        outs[threads] = (open(out_dir + "/pred.vcf").read(), open(out_dir + "/none.vcf").read())
    assert outs[4][0].count("\n") > 2048
    assert outs[1] == outs[4]
