"""CPU: the C++ candidate-TSV decoder (csrc/nsc_tsv.cpp) against the oracle restatement of
dataloader.py:38-58 / 267-274 on reference-format files (generate_dataset.py:1569-1581)."""
import os

import numpy as np
import pytest

from oracle import nsnet_oracle as oracle
from neusomatic_b200 import anns_to_255, synth


@pytest.fixture(scope="module")
def built():
    from neusomatic_b200 import build
    build.build_lib()


def _write(path, cand, with_anns):
    with open(path, "w") as f:
        for i in range(len(cand)):
            f.write(oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i], cand.anns[i] if with_anns else None))


@pytest.mark.parametrize("ensemble", [False, True])
def test_decode_matches_oracle(tmp_path, built, ensemble):
    from neusomatic_b200.tsv import CandidateTSV
    cand = synth.structured_pileups(300, seed=9, ensemble=ensemble)
    p = str(tmp_path / "candidates_0.tsv")
    _write(p, cand, ensemble)
    tsv = CandidateTSV(p, num_anns=93 if ensemble else 0, num_threads=4)
    assert len(tsv) == 300
    got = tsv.decode()
    lines = open(p).read().splitlines()
    for i in (0, 1, 150, 299):
        tag, im, anns = oracle.decode_tsv_line(lines[i])
        assert got.tags[i] == tag == cand.tags[i]
        np.testing.assert_array_equal(got.matrices[i], im)
        vt, center, length, tc, nc = oracle.parse_tag(tag)
        assert (got.center[i], got.tumor_cov[i], got.normal_cov[i]) == (center, tc, nc)
        assert oracle.VARTYPE_CLASSES[got.labels[i, 0]] == vt and got.labels[i, 1] == min(length, 3)
        if ensemble:
            np.testing.assert_array_equal(got.anns255[i], anns_to_255(np.float64(anns)))
    np.testing.assert_array_equal(got.matrices, cand.matrices)
    np.testing.assert_array_equal(got.center, cand.center)
    # ranges and single-thread decode give the same bytes
    part = CandidateTSV(p, num_anns=93 if ensemble else 0, num_threads=1).decode(37, 100)
    np.testing.assert_array_equal(part.matrices, cand.matrices[37:137])
    assert part.tags == cand.tags[37:137]


def test_annotation_text_is_parsed_like_python_float(tmp_path, built):
    """dataloader.py:44-58: anns = float(text); matrix value = a * 255.0 in float64, then .float()."""
    from neusomatic_b200.tsv import CandidateTSV
    cand = synth.structured_pileups(2, seed=3, ensemble=True)
    texts = ["0.1234", "1e-3", "0.3333333333333333", "1", "0"] + ["0.5"] * 88
    p = str(tmp_path / "c.tsv")
    with open(p, "w") as f:
        for i in range(2):
            line = oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i], None).rstrip("\n")
            f.write(line + "\t" + "\t".join(texts) + "\n")
    got = CandidateTSV(p, num_anns=93).decode()
    ref = (np.array([float(t) for t in texts], np.float64) * 255.0).astype(np.float32)
    np.testing.assert_array_equal(got.anns255[0], ref)


def test_annotation_fast_decimal_path_equals_python_float(tmp_path, built):
    """The decoder parses plain decimals on an exact fast path (<= 15 significant digits: one correctly rounded division)
    and leaves the rest to strtod: both must give Python's float(text) bit for bit, for every shape of literal the
    ensemble TSVs can hold (generate_dataset.py:1570-1571 writes str(round(x, 4)), but any float literal is legal)."""
    from neusomatic_b200.tsv import CandidateTSV
    rng = np.random.default_rng(7)
    cand = synth.structured_pileups(40, seed=3, ensemble=True)
    p = str(tmp_path / "d.tsv")
    all_texts = []
    with open(p, "w") as f:
        for i in range(40):
            texts = []
            for _ in range(93):
                kind = rng.integers(0, 8)
                nd = int(rng.integers(1, 18))                       # up to 17 significant digits: beyond the fast path too
                digs = "".join(str(int(d)) for d in rng.integers(0, 10, size=nd))
                if kind == 0:
                    t = "0." + digs
                elif kind == 1:
                    t = digs[:1] + "." + digs[1:] if nd > 1 else digs
                elif kind == 2:
                    t = "0.000" + digs
                elif kind == 3:
                    t = str(round(float(rng.random()), 4))          # what the reference writes
                elif kind == 4:
                    t = "%.3e" % float(rng.random())                # exponent form -> strtod
                elif kind == 5:
                    t = "-0." + digs
                elif kind == 6:
                    t = digs + "." if nd < 6 else "." + digs        # "12." and ".5" are legal Python float literals
                else:
                    t = "0." + digs + "0" * int(rng.integers(0, 6))   # trailing zeros
                texts.append(t)
            all_texts.append(texts)
            line = oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i], None).rstrip("\n")
            f.write(line + "\t" + "\t".join(texts) + "\n")
    got = CandidateTSV(p, num_anns=93).decode()
    ref = (np.array([[float(t) for t in row] for row in all_texts], np.float64) * 255.0).astype(np.float32)
    np.testing.assert_array_equal(got.anns255, ref)


def test_bulk_tags_equal_per_candidate_tags(tmp_path, built):
    """nsc_tsv_tags (one call for a range) == nsc_tsv_tag candidate by candidate == the tag field the oracle wrote."""
    from neusomatic_b200.tsv import CandidateTSV
    cand = synth.structured_pileups(37, seed=9)
    p = str(tmp_path / "t.tsv")
    _write(p, cand, False)
    t = CandidateTSV(p)
    assert t.tags() == [t.tag(i) for i in range(len(t))] == list(cand.tags)
    assert t.tags(5, 0) == [] and t.tags(36, 1) == [cand.tags[36]] and t.tags(3, 4) == list(cand.tags[3:7])
    t.close()


def test_empty_and_malformed_files(tmp_path, built):
    from neusomatic_b200.tsv import CandidateTSV
    from neusomatic_b200._lib import NscError
    p = str(tmp_path / "empty.tsv")
    open(p, "w").close()
    assert len(CandidateTSV(p)) == 0
    cand = synth.structured_pileups(3, seed=4)
    good = [oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i]) for i in range(3)]
    for bad in ("0\t1\tonly.three.fields\n", good[1].replace(cand.tags[1], "21.5.A.C.SNP.7.1.30"),   # 8 tag fields
                good[1][:-20] + "\n"):                                                                   # truncated image
        q = str(tmp_path / "bad.tsv")
        with open(q, "w") as f:
            f.write(good[0] + bad + good[2])
        t = CandidateTSV(q)
        with pytest.raises(NscError):
            t.decode()
    with pytest.raises(NscError):
        CandidateTSV(str(tmp_path / "missing.tsv"))


def test_tags_lens_one_pass_equals_the_two_calls(tmp_path, built):
    from neusomatic_b200.tsv import CandidateTSV
    cand = synth.structured_pileups(64, seed=10)
    p = str(tmp_path / "t.tsv")
    with open(p, "w") as f:
        for i in range(64):
            tag = ("work/dataset/" if i % 3 == 0 else "") + cand.tags[i]            # path_.split("/")[-1], call.py:88
            f.write(oracle.encode_tsv_line(i, tag, cand.matrices[i]))
    t = CandidateTSV(p)
    tags, lens = t.tags_lens()
    assert tags == list(cand.tags)
    np.testing.assert_array_equal(lens, t.allele_lens())
    np.testing.assert_array_equal(lens, [[len(x.split(".")[2]), len(x.split(".")[3])] for x in cand.tags])
    tags, lens = t.tags_lens(7, 5)
    assert tags == list(cand.tags[7:12]) and lens.shape == (5, 2)
    assert t.tags_lens(3, 0)[0] == []
    t.close()


def test_line_index_across_pieces_and_staging(tmp_path, built):
    """The line index is built piecewise by threads (32 MB pieces): a file of several pieces with blank lines, CRLF endings
    and no final newline indexes like a sequential scan; nsc_tsv_stage copies the lines verbatim and its spans delimit the
    image fields; nsc_tsv_decode_rows returns the listed candidates' matrices."""
    import base64
    import zlib
    from neusomatic_b200.tsv import CandidateTSV
    cand = synth.structured_pileups(256, seed=12)
    lines = [oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i]) for i in range(256)]
    n = 70000                                                   # ~ 90 MB: three pieces
    p = str(tmp_path / "big.tsv")
    with open(p, "w", newline="") as f:
        for i in range(n):
            ln = lines[i % 256]
            if i % 1000 == 17:
                f.write("\n   \t\n")                            # blank lines are skipped (dataloader.py: `if not line.strip()`)
            if i % 5000 == 3:
                ln = ln[:-1] + "\r\n"
            if i == n - 1:
                ln = ln[:-1]                                    # no newline at the end of the file
            f.write(ln)
    t = CandidateTSV(p, num_threads=4)
    assert len(t) == n
    for first in (0, 24000, 26000, 50000, n - 300):             # around the piece boundaries and at both ends
        got = t.decode(first, 300)
        want = cand.matrices[(first + np.arange(300)) % 256]
        np.testing.assert_array_equal(got.matrices, want)
        assert got.tags == [cand.tags[(first + i) % 256] for i in range(300)]
    first, m = 4990, 1200
    text = np.zeros(t.range_bytes(first, m) + 16, np.uint8)
    spans, meta = np.zeros((m, 2), np.uint32), np.zeros((m, 3), np.int32)
    t.stage(first, m, text, spans, meta)
    raw = open(p, "rb").read()
    start = raw.index(lines[first % 256].encode()[:40])
    assert text[:t.range_bytes(first, m)].tobytes() in raw
    for k in (0, 1, 13, 555, m - 1):
        o, ln = int(spans[k, 0]), int(spans[k, 1])
        mat = np.frombuffer(zlib.decompress(base64.b64decode(text[o:o + ln].tobytes())), np.uint8).reshape(5, 32, 23)
        np.testing.assert_array_equal(mat, cand.matrices[(first + k) % 256])
        i = (first + k) % 256
        assert list(meta[k]) == [cand.center[i], cand.tumor_cov[i], cand.normal_cov[i]]
    tags, lens, hashes = t.stage(first, m, text, spans, meta, with_tags=True)     # the same pass gathers tags, allele lengths, tag hashes
    t2, l2 = t.tags_lens(first, m)
    assert tags == t2 == [cand.tags[(first + k) % 256] for k in range(m)]
    np.testing.assert_array_equal(lens, l2)
    def fnv(txt):
        h = 14695981039346656037
        for c in txt.encode():
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    assert [int(h) for h in hashes[:300]] == [fnv(x) for x in tags[:300]]
    assert len(set(hashes[:256].tolist())) == 256 and hashes[0] == hashes[256]     # distinct tags, and the repeat 256 lines later
    rows = np.array([5, 69999, 25000, 5, 31234], np.int64)
    np.testing.assert_array_equal(t.decode_rows(rows), cand.matrices[rows % 256])
    with pytest.raises(Exception):
        t.decode_rows([n])
    with pytest.raises(Exception):
        t.stage(0, 10, np.zeros(100, np.uint8), spans, meta)   # text buffer too small
    t.close()
