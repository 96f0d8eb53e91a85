"""Dictionaries -> VCF stage (neusomatic_b200/vcf.py) against the outputs of the UNMODIFIED reference call.py
(tests/golden/vcf_*.npz, written by tests/golden/make_golden_vcf.py): record builders, none.vcf filter, sorting,
de-duplication, QUAL/SCORE formatting -- and the reference's whole call_neusomatic output on synthetic candidates,
reproduced from the oracle's forward pass."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from neusomatic_b200 import vcf
from neusomatic_b200.fasta import FastaFile, MemoryFasta
from oracle import nsnet_oracle as oracle, torch_port

CASES = ["std_v014", "std_wex", "ens_v014"]


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, "vcf_%s.npz" % name), allow_pickle=False))
    g["fasta"] = json.loads(str(g["fasta"]))
    return g


def dict_from(g, prefix):
    """Rebuilds a call_variants dictionary with the reference's element types (np.float32 scalars, [1]-shaped position)."""
    d = {}
    for i, path in enumerate(g[prefix + "_paths"]):
        d[str(path)] = [str(g[prefix + "_type"][i]), np.array([g[prefix + "_pos"][i]], np.float32), np.int64(g[prefix + "_len"][i]),
                        list(g[prefix + "_p1"][i]), list(g[prefix + "_p3"][i]), list(g[prefix + "_o1"][i]), list(g[prefix + "_o3"][i])]
    return d


def images(g):
    nt = oracle.non_transformed(g["matrices"], g["center"])
    return {str(t): nt[i] for i, t in enumerate(g["tags"])}


def run_stage(g, final_preds, none_preds):
    chroms = [str(c) for c in g["chroms"]]
    order = {c: i for i, c in enumerate(chroms)}
    fa = MemoryFasta(g["fasta"])
    img = images(g)
    recs = dict(vcf.pred_vcf_records(fa, final_preds, img, chroms, oracle.VARTYPE_CLASSES, 1))
    pred = "".join(vcf.vcf_lines(vcf.get_vcf_records(recs), order, 0.7, 0.4))
    none = "".join(vcf.vcf_lines(vcf.get_vcf_records(dict(vcf.pred_vcf_records_none(none_preds, chroms))), order, 0.7, 0.4))
    return pred, none


@pytest.mark.parametrize("name", CASES)
def test_forced_dictionaries_give_the_reference_lines(name):
    g = load(name)
    forced = dict_from(g, "forced")
    pred, none = run_stage(g, forced, forced)
    assert pred == str(g["forced_vcf"])
    assert none == str(g["forced_none_vcf"])
    assert pred.count("\n") > 50          # the builders did produce records (SNP, INS and DEL rows)
    kinds = {(len(l.split("\t")[3]) > 1, len(l.split("\t")[4]) > 1) for l in pred.splitlines()}
    assert kinds >= {(False, False), (True, False), (False, True)}      # (True, True): multi-base SNP records


@pytest.mark.parametrize("name", CASES)
def test_reference_dictionaries_give_the_reference_vcf(name):
    g = load(name)
    pred, none = run_stage(g, dict_from(g, "final"), dict_from(g, "none"))
    assert pred == str(g["pred_vcf"])
    assert none == str(g["none_vcf"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_forward_reproduces_the_reference_call(name):
    """oracle forward -> oracle call_variants -> this VCF stage == pred.vcf / none.vcf of the reference's call_neusomatic."""
    g = load(name)
    ck = torch.load(os.path.join(GOLDEN, name + ".pth"), map_location="cpu", weights_only=True)
    x = oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], g.get("anns"), int(ck["coverage_thr"]))
    o1, o2, o3 = torch_port.forward(torch_port.prepare(ck["state_dict"]), torch.from_numpy(x))
    final_preds, none_preds = oracle.call_variants(np.asarray(o1), np.asarray(o2), np.asarray(o3), [str(t) for t in g["tags"]])
    ref_final, ref_none = dict_from(g, "final"), dict_from(g, "none")
    assert set(final_preds) == set(ref_final) and set(none_preds) == set(ref_none)
    pred, none = run_stage(g, final_preds, none_preds)
    assert pred == str(g["pred_vcf"])
    assert none == str(g["none_vcf"])


def test_write_vcf_and_prob2phred(tmp_path):
    recs = [["chr2", 5, "A", "T", np.float32(0.9997) * np.float32(0.9998)], ["chr1", 9, "G", "GA", np.float32(1.0)],
            ["chr1", 9, "G", "GA", np.float32(1.0)], ["chr1", 3, "C", "C", np.float32(0.9)], ["chr1", 2, "T", "A", np.float32(0.45)],
            ["chr1", 1, "T", "G", np.float32(0.0)], []]
    out = tmp_path / "o.vcf"
    vcf.write_vcf(recs, str(out), {"chr1": 0, "chr2": 1}, 0.7, 0.4)
    lines = out.read_text().splitlines()
    assert lines == ["chr1\t1\t.\tT\tG\t0.0000\tREJECT\tSCORE=0.0000\tGT\t0/1",
                     "chr1\t2\t.\tT\tA\t2.5964\tLowQual\tSCORE=0.4500\tGT\t0/1",
                     "chr1\t9\t.\tG\tGA\t100.0000\tPASS\tSCORE=1.0000\tGT\t0/1",
                     "chr2\t5\t.\tA\tT\t33.0111\tPASS\tSCORE=0.9995\tGT\t0/1"]     # 33.0111: test/NeuSomatic_standalone.vcf:17
    with pytest.raises(AssertionError):
        vcf.prob2phred(1.5)
    assert vcf.get_type("AC", "A") == "DEL" and vcf.get_type("A", "ACT,G") == "INS" and vcf.get_type("A", "G") == "SNP"


def test_fasta_reader(tmp_path):
    seqs = {"chr1": "ACGTACGTACGTAAACCCGGGTTT" * 7, "chrM": "GATTACA"}
    p = tmp_path / "r.fa"
    with open(p, "w") as f:
        for k, s in seqs.items():
            f.write(">%s some description\n" % k)
            for i in range(0, len(s), 50):
                f.write(s[i:i + 50] + "\n")
    with FastaFile(str(p)) as fa:
        assert fa.references == ("chr1", "chrM") and fa.lengths == (168, 7)
        for a, b in ((0, 1), (0, 50), (49, 51), (3, 160), (100, 168), (160, 400), (5, 5)):
            assert fa.fetch("chr1", a, b) == seqs["chr1"][a:b]
        assert fa.fetch("chrM", 2, 6) == "TTAC"
    # a .fai written by samtools has the same five columns
    with open(str(p) + ".fai", "w") as f:
        f.write("chr1\t168\t23\t50\t51\nchrM\t7\t%d\t7\t8\n" % (23 + 168 + 4 + 23))
    with FastaFile(str(p)) as fa:
        assert fa.fetch("chr1", 49, 120) == seqs["chr1"][49:120] and fa.fetch("chrM", 0, 7) == "GATTACA"


def _planted_tables(n, seed, groups):
    """Genome-consistent candidates with planted (noisy) predictions, the called ones dealt into ``groups`` PredTables:
    ([(PredTable, images)], chroms, seqs)."""
    from neusomatic_b200 import synth
    from neusomatic_b200.call import PredTable, non_transformed_images
    cand, seqs, chroms = synth.genome_candidates(n, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    a1 = np.array([oracle.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags])
    ln = np.array([min(int(t.split(".")[6]), 3) for t in cand.tags])
    p1 = rng.dirichlet([0.3] * 4, size=n).astype(np.float32)
    p1[np.arange(n), a1] += 2
    p1 /= p1.sum(1, keepdims=True)
    p3 = rng.dirichlet([0.3] * 4, size=n).astype(np.float32)
    p3[np.arange(n), ln] += 2
    p3 /= p3.sum(1, keepdims=True)
    o2 = (np.asarray(cand.center, np.float32) + rng.normal(0, 0.6, n).astype(np.float32)).reshape(-1, 1)
    called = np.nonzero(a1 != 2)[0]
    out = []
    for keep in np.array_split(called, groups):
        tab = PredTable()
        for lo in range(0, len(keep), 700):                  # several parts, as the pipeline files them
            k = keep[lo:lo + 700]
            tab.add([cand.tags[i] for i in k], a1[k], o2[k], ln[k], p1[k], p3[k], np.log(p1[k]), np.log(p3[k]))
        img = non_transformed_images(cand.matrices[keep], np.asarray(cand.center)[keep])
        out.append((tab, {cand.tags[i]: im for i, im in zip(keep, img)}))
    return out, chroms, seqs


def _same(x, y):
    if isinstance(x, (list, tuple)):
        return len(x) == len(y) and all(_same(u, v) for u, v in zip(x, y))
    if isinstance(x, np.ndarray):
        return x.dtype == y.dtype and np.array_equal(x, y)
    return type(x) is type(y) and x == y


def test_record_pool_equals_the_serial_stage(tmp_path, monkeypatch):
    """call.py:308-333 builds the records in a process pool; here the pool gets chunks of the table as arrays and works
    while the caller scores the next group.  Same records, same order, same element types as the in-process loop -- also
    for a FASTA without .fai (the parent's index travels with the jobs) and with the [path, pred] tail left out."""
    tabs, chroms, seqs = _planted_tables(5200, 3, 2)
    fa = str(tmp_path / "genome.fa")
    with open(fa, "w") as f:
        for c in chroms:
            f.write(">%s\n" % c)
            f.writelines(seqs[c][i:i + 60] + "\n" for i in range(0, len(seqs[c]), 60))
    monkeypatch.setattr(vcf, "POOL_MIN", 1000)
    serial = [vcf.pred_vcf_records(MemoryFasta(seqs), t[0], t[1], chroms, oracle.VARTYPE_CLASSES, 1) for t in tabs]
    assert all(len(s) > 800 for s in serial)
    monkeypatch.setattr(vcf, "POOL_MIN", 250)                # (the one-shot form starts workers from 4 x POOL_MIN records)
    one_shot = vcf.pred_vcf_records(fa, tabs[0][0], tabs[0][1], chroms, oracle.VARTYPE_CLASSES, 3)
    assert _same(one_shot, serial[0])
    monkeypatch.setattr(vcf, "POOL_MIN", 1000)
    with vcf.RecordPool(fa, chroms, oracle.VARTYPE_CLASSES, 3, with_pred=False) as rp:
        handles = [rp.submit(t[0], t[1]) for t in tabs]      # the second is submitted while the first is being built
        assert handles[0][0] == "pending" and rp.pool is not None
        got = [rp.collect(h) for h in handles]
    for g, s in zip(got, serial):
        assert len(g) == len(s)
        assert all(x[0] == y[0] and _same(x[1][:5], y[1][:5]) and x[1][5] is None for x, y in zip(g, s))
    lines = lambda recs: "".join(vcf.vcf_lines(vcf.get_vcf_records(dict(recs)), {c: i for i, c in enumerate(chroms)}, 0.7, 0.4))
    assert lines(got[0]) == lines(serial[0]) and lines(got[1]) == lines(serial[1])


def test_vcf_lines_array_path_equals_the_record_by_record_path():
    """write_vcf's lines (call.py:366-387): the whole-array rounding / formatting path against the reference's expressions
    applied record by record, on 6 x 10^4 records including the exact thresholds, 0, 1 and probabilities within 1e-12 of 1."""
    rng = np.random.Generator(np.random.PCG64(0))
    n = 60000
    p = rng.random(n).astype(np.float32)
    p[:200], p[200:400] = np.float32(0), np.float32(1)
    p[400:2000] = (1 - 10 ** (-rng.uniform(3, 12, 1600))).astype(np.float32)
    p[2000:4000] = rng.choice(np.array([0.7, 0.4, 0.69995, 0.70005, 0.39995, 0.40005], np.float32), 2000)
    recs = [["chr%d" % (i % 3), int(rng.integers(1, 10 ** 5)), "A", "AC"[i % 2], p[i]] for i in range(n)]
    order = {"chr0": 0, "chr1": 1, "chr2": 2}
    srt = [r for r in sorted(recs, key=lambda x: [order[x[0]], x[1]]) if r[2] != r[3]]
    want = vcf._vcf_lines_scalar(srt, 0.7, 0.4)
    assert vcf.vcf_lines(recs, order, 0.7, 0.4) == want and len(want) > 25000
    mixed = [list(r) for r in recs[:500]]
    mixed[7][4] = float(mixed[7][4])                         # one Python float: the whole call takes the scalar path
    srt = [r for r in sorted(mixed, key=lambda x: [order[x[0]], x[1]]) if r[2] != r[3]]
    assert vcf.vcf_lines(mixed, order, 0.7, 0.4) == vcf._vcf_lines_scalar(srt, 0.7, 0.4)


def test_record_pool_propagates_a_failed_record(tmp_path, monkeypatch):
    """call.py:326-329: a worker that fails returns None and the stage raises -- in the pool as in the calling process."""
    tabs, chroms, seqs = _planted_tables(2600, 6, 1)
    tab, images = tabs[0]
    fa = str(tmp_path / "genome.fa")
    with open(fa, "w") as f:
        for c in chroms:
            f.write(">%s\n" % c)
            f.writelines(seqs[c][i:i + 60] + "\n" for i in range(0, len(seqs[c]), 60))
    victim = list(tab.keys())[1500]
    images[victim] = str(tmp_path / "no_such_image.png")      # the reference's PNG hand-over with a missing file: raises inside
    monkeypatch.setattr(vcf, "POOL_MIN", 250)
    for threads in (1, 3):
        with pytest.raises(Exception, match="pred_vcf_records_path failed"):
            vcf.pred_vcf_records(fa, tab, {k: (v if k != victim else images[victim]) for k, v in images.items()}, chroms,
                                 oracle.VARTYPE_CLASSES, threads)
