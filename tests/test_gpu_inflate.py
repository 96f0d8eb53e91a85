"""The device decoder of the TSV's image fields (csrc/nsc_inflate_gpu.cu: base64 + zlib inflate, one warp per candidate)
against the reference's extract_zlib (dataloader.py:38-39, ``zlib.decompress(base64.b64decode(field))``): bit-exact on
every stream it accepts, never wrong bytes on malformed input (it declines and the host's zlib decides), and the whole
``nsc_score_host_b64`` / ``call_tsv`` path gives the bytes of the host-inflate path."""
import base64
import os
import zlib

import numpy as np
import pytest
import torch

from neusomatic_b200 import Engine, _lib, synth
from neusomatic_b200.engine import anns_to_255
from oracle import nsnet_oracle as oracle

pytestmark = pytest.mark.gpu
N = 5 * 32 * 23


def device_inflate(fields, offsets_pad=(0, 1, 2, 3, 5, 7, 16)):
    """fields: list of bytes (base64 text).  Packs them at odd offsets with junk in between; returns (matrices, declined)."""
    lib = _lib.load()
    parts, spans, pos = [], [], 0
    for i, f in enumerate(fields):
        pad = offsets_pad[i % len(offsets_pad)]
        parts.append(b"\t" * pad)
        pos += pad
        spans.append((pos, len(f)))
        parts.append(f)
        pos += len(f)
    text = np.frombuffer(b"".join(parts) + b"\n" * 16, np.uint8).copy()
    sp = np.asarray(spans, np.uint32).reshape(-1, 2)
    out = np.empty((len(fields), N), np.uint8)
    decl = np.empty(len(fields), np.int32)
    _lib.check(lib.nsc_debug_inflate_b64(0, text.ctypes.data, text.size, sp.ctypes.data, len(fields), out.ctypes.data, decl.ctypes.data))
    return out, decl


def streams_of(data):
    for level in (0, 1, 6, 9):
        yield zlib.compress(data, level)
    for strategy in (zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
        c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strategy)
        yield c.compress(data) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, 9)          # small window
    yield c.compress(data) + c.flush()
    c = zlib.compressobj(6)                              # several blocks (sync flushes, an empty stored block each)
    parts = [c.compress(data[i:i + 997]) + c.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(data), 997)]
    yield b"".join(parts) + c.flush()


def matrices_of_many_shapes():
    rng = np.random.Generator(np.random.PCG64(12))
    cand = synth.structured_pileups(60, seed=8)
    datas = [cand.matrices[i].tobytes() for i in range(60)]
    datas += [synth.uniform_pileups(8, seed=3).matrices[i].tobytes() for i in range(8)]
    datas += [synth.genome_candidates(16, seed=5, ensemble=False)[0].matrices[i].tobytes() for i in range(16)]
    for k in range(6):
        datas.append(rng.integers(0, 256, N, dtype=np.uint8).tobytes())                                            # incompressible
        datas.append((rng.integers(0, 4, N, dtype=np.uint8) * (rng.random(N) < 0.2)).astype(np.uint8).tobytes())   # sparse
        datas.append((rng.integers(0, 256, N, dtype=np.uint8) * (rng.random(N) < 0.02 * (k + 1))).astype(np.uint8).tobytes())
    datas.append(bytes(N))                                                                                         # one long run
    datas.append((np.arange(N) % 23).astype(np.uint8).tobytes())                                                   # channel stride
    datas.append((np.arange(N) % 3).astype(np.uint8).tobytes())                                                    # distance-3 matches
    datas.append((np.arange(N) % 251).astype(np.uint8).tobytes())
    datas.append(np.repeat(rng.integers(0, 256, N // 40, dtype=np.uint8), 40).tobytes())                           # runs of 40
    # geometric symbol frequencies (1, 1, 2, 4, ... 1024 occurrences): Huffman codes up to 12-13 bits, longer than the decoder's
    # 10-bit look-up table -> the bit-by-bit canonical walk; the matching distances of the shuffled copy do the same to the
    # 8-bit distance table
    geo = np.concatenate([np.full(max(1, 1 << max(k - 1, 0)), 200 - k, np.uint8) for k in range(12)])
    geo = np.concatenate([geo, np.full(N - geo.size, 7, np.uint8)])
    for seed in (1, 2, 3):
        datas.append(np.random.Generator(np.random.PCG64(seed)).permutation(geo).tobytes())
    far = rng.integers(0, 256, N, dtype=np.uint8)
    for k in range(0, N - 40, 97):                                                                                 # matches at many distinct distances
        d = int(rng.integers(1, k + 1)) if k else 1
        if k >= d + 8:
            far[k:k + 8] = far[k - d:k - d + 8]
    datas.append(far.tobytes())
    return datas


def test_device_inflate_equals_zlib_on_many_shapes():
    datas = matrices_of_many_shapes()
    fields, want = [], []
    for d in datas:
        for s in streams_of(d):
            assert zlib.decompress(s) == d
            fields.append(base64.b64encode(s))
            want.append(np.frombuffer(d, np.uint8))
    got, decl = device_inflate(fields)
    assert int(decl.sum()) == 0, "declined %d of %d valid streams (first: %d)" % (decl.sum(), len(fields), int(np.nonzero(decl)[0][0]))
    np.testing.assert_array_equal(got, np.stack(want))


def test_device_inflate_never_returns_wrong_bytes():
    rng = np.random.Generator(np.random.PCG64(5))
    cand = synth.structured_pileups(12, seed=9)
    fields, expect = [], []          # expect: bytes the reference would return, or None when it raises
    def ref(field):
        try:
            d = zlib.decompress(base64.b64decode(field, validate=True))
            return d if len(d) == N else None
        except Exception:
            return None
    for i in range(12):
        d = cand.matrices[i].tobytes()
        s = zlib.compress(d)
        cases = [s[:cut] for cut in (1, 2, 5, 12, len(s) // 2, len(s) - 5, len(s) - 1)]                    # truncated
        cases += [s + b"\0", s + s]                                                                         # trailing bytes
        cases += [zlib.compress(d[:-1]), zlib.compress(d + b"\1")]                                          # wrong size
        cases.append(s[:-1] + bytes([s[-1] ^ 1]))                                                           # wrong Adler-32
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_DEFAULT_STRATEGY, d[:64])                     # preset dictionary
        cases.append(co.compress(d) + co.flush())
        cf = zlib.compressobj(6)                                                                            # a valid stream of > 8192 characters
        cases.append(b"".join(cf.compress(d[k:k + 4]) + cf.flush(zlib.Z_SYNC_FLUSH) for k in range(0, len(d), 4)) + cf.flush())
        assert len(base64.b64encode(cases[-1])) > 8192 and zlib.decompress(cases[-1]) == d
        for _ in range(120):                                                                                # bit flips
            b = bytearray(s)
            for _k in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            cases.append(bytes(b))
        for c in cases:
            fields.append(base64.b64encode(c))
            expect.append(ref(fields[-1]))
        bad = bytearray(base64.b64encode(s))
        bad[len(bad) // 3] = ord("*")                                                                       # not a base64 character
        fields.append(bytes(bad))
        expect.append(None)
        fields.append(base64.b64encode(s).rstrip(b"="))                                                     # padding stripped: same bytes
        expect.append(d)
    got, decl = device_inflate(fields)
    accepted = wrong = 0
    for i, e in enumerate(expect):
        if decl[i]:
            assert not got[i].any()              # a declined candidate's row is never written
            continue
        accepted += 1
        if e is None or got[i].tobytes() != e:
            wrong += 1
    assert wrong == 0, "%d accepted streams with bytes the reference would not return" % wrong
    assert accepted >= 12 and int(decl.sum()) > 1000


def _staged(cand, anns=None):
    """Reference-format lines of the candidates in one text buffer + spans of their image fields (what nsc_tsv_stage makes)."""
    lines = [oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i], None if anns is None else anns[i]).encode() for i in range(len(cand.tags))]
    spans, pos = [], 0
    for ln in lines:
        f = ln.split(b"\t")
        off = pos + len(f[0]) + len(f[1]) + len(f[2]) + 3
        spans.append((off, len(f[3].rstrip(b"\n"))))
        pos += len(ln)
    return bytearray(b"".join(lines) + b"\0" * 64), np.asarray(spans, np.uint32)


@pytest.mark.parametrize("ensemble", [False, True])
def test_score_host_b64_is_bit_equal_to_score_host_u8(ensemble):
    C = 119 if ensemble else 26
    n = 20000 + 37
    cand = synth.structured_pileups(n, seed=11, ensemble=ensemble)
    sd = synth.random_state_dict(C, seed=7)
    eng = Engine(sd, C)
    meta = np.stack([cand.center, cand.tumor_cov, cand.normal_cov], 1).astype(np.int32)
    a255 = anns_to_255(cand.anns) if ensemble else None
    want = [t.numpy().copy() for t in eng.score_host_u8(cand.matrices, meta, a255, 100.0)]
    text, spans = _staged(cand)
    # three candidates the device decoder must hand to the host: trailing bytes after the stream (zlib.decompress accepts them)
    odd = [5, 4096 + 17, n - 1]
    parts, sp2, pos = [], spans.copy(), 0
    raw = bytes(text)
    extra = []
    for k in odd:
        s = zlib.compress(np.ascontiguousarray(cand.matrices[k]).tobytes()) + b"\0\0\0"
        extra.append(base64.b64encode(s))
    base = len(raw)
    for k, f in zip(odd, extra):
        sp2[k] = (base, len(f))
        base += len(f) + 3
    raw = raw + b"\t\t\t".join(extra) + b"\0" * 64
    tbuf = torch.from_numpy(np.frombuffer(raw, np.uint8).copy()).pin_memory()
    got, nhost = eng.score_host_b64(tbuf, torch.from_numpy(sp2.view(np.int32)), meta, a255, 100.0)
    assert nhost == 3
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g.numpy(), w)
    # a field zlib rejects fails the call, with the candidate named
    bad = bytearray(raw)
    o, ln = int(sp2[1234][0]), int(sp2[1234][1])
    bad[o + ln // 2] = ord("A") if bad[o + ln // 2] != ord("A") else ord("B")
    bad[o + ln // 2 + 1] = ord("A") if bad[o + ln // 2 + 1] != ord("A") else ord("B")
    with pytest.raises(RuntimeError, match="candidate 1234"):
        eng.score_host_b64(np.frombuffer(bytes(bad), np.uint8).copy(), sp2.view(np.int32), meta, a255, 100.0)
    # ... and the handle still works afterwards
    got, _ = eng.score_host_b64(tbuf, torch.from_numpy(sp2.view(np.int32)), meta, a255, 100.0)
    np.testing.assert_array_equal(got[0].numpy(), want[0])
    eng.close()


@pytest.mark.parametrize("ensemble", [False, True])
def test_call_tsv_device_inflate_equals_host_inflate(tmp_path, monkeypatch, ensemble):
    from neusomatic_b200.call import CandidateScorer
    n = 9000
    cand = synth.structured_pileups(n, seed=23, ensemble=ensemble)
    p = str(tmp_path / "candidates.tsv")
    with open(p, "w") as f:
        for i in range(n):
            f.write(oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i], cand.anns[i] if ensemble else None))
    C = 119 if ensemble else 26
    ck = {"state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in synth.random_state_dict(C, seed=3).items()},
          "coverage_thr": 100, "normalize_channels": False, "tag": "t", "epoch": 0}
    sc = CandidateScorer(ck, ensemble=ensemble)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NSC_TSV_HOST_DECODE", mode)
        tp = {}
        fp, npred = sc.call_tsv(p, batch=4096, num_threads=4, true_path=tp)
        res[mode] = (dict(fp), dict(npred), tp)
    (f1, n1, t1), (f0, n0, t0) = res["1"], res["0"]
    assert len(f1) + len(n1) == n and list(f1) == list(f0) and list(n1) == list(n0) and list(t1) == list(t0)
    for a, b in ((f1, f0), (n1, n0)):
        for k in a:
            for x, y in zip(a[k], b[k]):
                np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    for k in t1:
        np.testing.assert_array_equal(t1[k], t0[k])
    sc.engine.close()


def test_score_host_b64_argument_errors():
    import ctypes as C
    from neusomatic_b200._lib import NscError
    lib = _lib.load()
    cand = synth.structured_pileups(64, seed=2)
    eng = Engine(synth.random_state_dict(26, seed=1), 26)
    text, spans = _staged(cand)
    text = np.frombuffer(bytes(text), np.uint8).copy()
    meta = np.stack([cand.center, cand.tumor_cov, cand.normal_cov], 1).astype(np.int32)
    outs, nhost = eng.score_host_b64(text, spans[:0].view(np.int32), meta[:0], None, 100.0)          # empty batch
    assert outs[0].shape == (0, 4) and nhost == 0
    bad = spans.copy()
    bad[10, 0] = text.size                                                                              # span outside the text
    with pytest.raises(NscError, match="span outside"):
        eng.score_host_b64(text, bad.view(np.int32), meta, None, 100.0)
    o = Engine._host_out(64, True, False)
    args = [eng.h, text.ctypes.data, text.size, None, meta.ctypes.data, None, C.c_float(100.0), 64] + [_lib.ptr(t) for t in o] + [None]
    assert lib.nsc_score_host_b64(*args) != 0 and b"spans" in lib.nsc_last_error()                      # NULL spans
    args[3], args[4] = spans.ctypes.data, None
    assert lib.nsc_score_host_b64(*args) != 0                                                           # NULL meta
    args[4], args[6] = meta.ctypes.data, C.c_float(0.0)
    assert lib.nsc_score_host_b64(*args) != 0                                                           # coverage_thr must be positive
    ens = Engine(synth.random_state_dict(119, seed=1), 119)
    with pytest.raises(NscError, match="anns255"):
        ens.score_host_b64(text, spans.view(np.int32), meta, None, 100.0)                               # the ensemble needs annotations
    ens.close()
    got, _ = eng.score_host_b64(text, spans.view(np.int32), meta, None, 100.0)                          # and the handle is intact
    want = eng.score_host_u8(cand.matrices, meta, None, 100.0)
    np.testing.assert_array_equal(got[0].numpy(), want[0].numpy())
    eng.close()
