"""The TSV decoder's own zlib-stream inflate (csrc/nsc_inflate.h) against zlib: whenever it accepts a stream its output is
zlib's, byte for byte; it may decline a stream (the decoder then falls back to zlib) but must never return wrong bytes --
including on truncated and corrupted input."""
import ctypes as C
import zlib

import numpy as np
import pytest

from neusomatic_b200 import _lib, build as nsc_build, synth


@pytest.fixture(scope="module")
def lib():
    nsc_build.build_lib()
    return _lib.load()


def fast(lib, stream, n):
    src = np.frombuffer(stream, np.uint8).copy() if len(stream) else np.zeros(1, np.uint8)
    out = np.empty(max(n, 1), np.uint8)
    rc = lib.nsc_debug_inflate(src.ctypes.data, len(stream), out.ctypes.data, n, 0)
    return rc, out[:n].tobytes()


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


def test_fast_inflate_equals_zlib_on_many_shapes(lib):
    rng = np.random.Generator(np.random.PCG64(12))
    cand = synth.structured_pileups(40, seed=8)
    datas = [cand.matrices[i].tobytes() for i in range(40)]
    datas += [synth.uniform_pileups(4, seed=3).matrices[i].tobytes() for i in range(4)]
    for n in (0, 1, 2, 7, 100, 257, 258, 259, 3680, 70000):
        datas.append(rng.integers(0, 256, n, dtype=np.uint8).tobytes())                       # incompressible
        datas.append((rng.integers(0, 4, n, dtype=np.uint8) * (rng.random(n) < 0.2)).astype(np.uint8).tobytes())   # sparse
        datas.append(bytes(n))                                                                    # one long run
        datas.append((np.arange(n) % 23).astype(np.uint8).tobytes())                              # period 23 (channel stride)
        datas.append((np.arange(n) % 3).astype(np.uint8).tobytes())                               # distance 3 matches
    taken = total = 0
    for d in datas:
        for s in streams_of(d):
            total += 1
            rc, out = fast(lib, s, len(d))
            assert rc in (0, 2), lib.nsc_last_error()
            if rc == 0:
                taken += 1
                assert out == d
            # the zlib path of the same hook always accepts a valid stream
            o2 = np.empty(max(len(d), 1), np.uint8)
            src = np.frombuffer(s, np.uint8).copy()
            assert lib.nsc_debug_inflate(src.ctypes.data, len(s), o2.ctypes.data, len(d), 1) == 0
            assert o2[:len(d)].tobytes() == d
    assert taken >= 0.95 * total, (taken, total)          # the fast path is the common path, not a curiosity


def test_fast_inflate_never_returns_wrong_bytes(lib):
    rng = np.random.Generator(np.random.PCG64(5))
    cand = synth.structured_pileups(12, seed=9)
    declined = 0
    for i in range(12):
        d = cand.matrices[i].tobytes()
        s = zlib.compress(d)
        for cut in (1, 2, 5, len(s) // 2, len(s) - 5, len(s) - 1):          # truncated
            rc, out = fast(lib, s[:cut], len(d))
            assert rc != 0
        rc, out = fast(lib, s + b"\\0", len(d))                              # trailing garbage: the fast path declines (zlib itself ignores bytes after the stream)
        assert rc != 0
        rc, out = fast(lib, s, len(d) - 1)                                   # wrong expected size
        assert rc != 0
        rc, out = fast(lib, s, len(d) + 1)
        assert rc != 0
        for _ in range(150):                                                 # bit flips anywhere in the stream
            b = bytearray(s)
            for _k in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            rc, out = fast(lib, bytes(b), len(d))
            if rc == 0:
                assert out == d or zlib.decompress(bytes(b)) == out          # accepted => exactly what zlib would give
            else:
                declined += 1
    assert declined > 1000


def test_decoder_uses_it_and_zlib_gives_the_same_batch(lib, tmp_path, monkeypatch):
    from neusomatic_b200.tsv import CandidateTSV
    from oracle import nsnet_oracle as oracle
    cand = synth.structured_pileups(500, seed=21)
    p = str(tmp_path / "c.tsv")
    with open(p, "w") as f:
        for i in range(500):
            f.write(oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i]))
    t = CandidateTSV(p, num_threads=3)
    got = t.decode()
    np.testing.assert_array_equal(got.matrices, cand.matrices)
    t.close()
