"""Candidate TSV reader on top of the C++ decoder (csrc/nsc_tsv.cpp): the inference-side replacement of
the reference's ``extract_info_tsv`` / ``candidate_loader_tsv`` / tag parsing (dataloader.py:38-123,
267-274).  Decodes ranges of candidates into (optionally pinned) batch buffers in the layout
``nsc_score_host_u8`` consumes."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .synth import Candidates


class CandidateTSV:
    def __init__(self, path, num_anns=0, num_threads=0):
        self.lib = _lib.load()
        self.num_anns = int(num_anns)
        self.num_threads = int(num_threads)
        h = C.c_void_p()
        _lib.check(self.lib.nsc_tsv_open(C.byref(h), str(path).encode()))
        self.h = h

    def __len__(self):
        return int(self.lib.nsc_tsv_count(self.h))

    def tag(self, i):
        p, n = C.c_char_p(), C.c_int32()
        _lib.check(self.lib.nsc_tsv_tag(self.h, i, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value).decode()

    def tags(self, first=0, n=None):
        """Tags of candidates [first, first + n) in one library call (the keys of call.py's dictionaries)."""
        n = len(self) - first if n is None else n
        if n == 0:
            return []
        need = C.c_int64()
        cap = 96 * n + 64                                  # one pass for ordinary tags (~40 bytes); sized exactly otherwise
        buf = np.empty(cap, np.uint8)
        rc = self.lib.nsc_tsv_tags(self.h, first, n, buf.ctypes.data, cap, C.byref(need))
        if rc != 0 and need.value > cap:
            buf = np.empty(need.value, np.uint8)
            rc = self.lib.nsc_tsv_tags(self.h, first, n, buf.ctypes.data, need.value, C.byref(need))
        _lib.check(rc)
        return buf[:need.value - 1].tobytes().decode().split("\n")

    def tags_lens(self, first=0, n=None):
        """(tags without their directory part, int32 [n,2] allele lengths) in one pass over the lines."""
        n = len(self) - first if n is None else n
        lens = np.empty((n, 2), np.int32)
        if n == 0:
            return [], lens
        need = C.c_int64()
        cap = 96 * n + 64
        buf = np.empty(cap, np.uint8)
        rc = self.lib.nsc_tsv_tags_lens(self.h, first, n, buf.ctypes.data, cap, C.byref(need), lens.ctypes.data)
        if rc != 0 and need.value > cap:
            buf = np.empty(need.value, np.uint8)
            rc = self.lib.nsc_tsv_tags_lens(self.h, first, n, buf.ctypes.data, need.value, C.byref(need), lens.ctypes.data)
        _lib.check(rc)
        return buf[:need.value - 1].tobytes().decode().split("\n"), lens

    def allele_lens(self, first=0, n=None):
        """int32 [n,2] = (len(ref), len(alt)) of the candidates' tags (what the none.vcf filter branches on, call.py:343-350)."""
        n = len(self) - first if n is None else n
        lens = np.empty((n, 2), np.int32)
        if n:
            _lib.check(self.lib.nsc_tsv_allele_lens(self.h, first, n, lens.ctypes.data))
        return lens

    def decode(self, first=0, n=None, out=None, with_tags=True):
        """Returns Candidates(matrices uint8 [n,5,32,23], center, tumor_cov, normal_cov, anns255-or-None, tags).
        ``out`` may hold preallocated (pinned) arrays: dict(matrices=, meta=, anns255=)."""
        n = len(self) - first if n is None else n
        out = out or {}
        m = out.get("matrices")
        if m is None:
            m = np.empty((n, 5, 32, 23), np.uint8)
        meta = out.get("meta")
        if meta is None:
            meta = np.empty((n, 3), np.int32)
        a255 = out.get("anns255")
        if a255 is None and self.num_anns:
            a255 = np.empty((n, self.num_anns), np.float32)
        labels = np.empty((n, 2), np.int32)
        _lib.check(self.lib.nsc_tsv_decode(self.h, first, n, self.num_anns, _lib.ptr(m), _lib.ptr(meta),
                                           _lib.ptr(a255) if self.num_anns else None, labels.ctypes.data, self.num_threads))
        tags = self.tags(first, n) if with_tags else []
        meta_np = np.asarray(meta)
        c = Candidates(m, meta_np[:, 0], meta_np[:, 1], meta_np[:, 2], None, tags)
        c.anns255 = a255
        c.labels = labels
        return c

    def range_bytes(self, first=0, n=None):
        """Bytes of the lines of candidates [first, first + n): the text a staged slice needs (+ 16 of slack)."""
        n = len(self) - first if n is None else n
        return int(self.lib.nsc_tsv_range_bytes(self.h, first, n))

    def stage(self, first, n, text, spans, meta, anns255=None, labels=None, with_tags=False):
        """Everything of ``decode`` except the inflate, for ``Engine.score_host_b64``: copies the lines verbatim into ``text``
        (uint8, >= range_bytes + 16), fills ``spans`` [n,2] uint32 (offset, characters of each image field), ``meta`` [n,3]
        and ``anns255``.  with_tags: also returns (tags, allele lengths [n,2], uint64 tag hashes [n]) gathered in the same pass."""
        cap = int(text.numel() if hasattr(text, "numel") else text.size)
        args = [self.h, first, n, self.num_anns, _lib.ptr(text), cap, _lib.ptr(spans), _lib.ptr(meta),
                _lib.ptr(anns255) if self.num_anns else None, _lib.ptr(labels), self.num_threads]
        if not with_tags or n == 0:
            _lib.check(self.lib.nsc_tsv_stage(*args, None, 0, None, None, None))
            return ([], np.empty((0, 2), np.int32), np.empty(0, np.uint64)) if with_tags else None
        lens = np.empty((n, 2), np.int32)
        hashes = np.empty(n, np.uint64)
        need = C.c_int64()
        tcap = 96 * n + 64
        buf = np.empty(tcap, np.uint8)
        rc = self.lib.nsc_tsv_stage(*args, buf.ctypes.data, tcap, C.byref(need), lens.ctypes.data, hashes.ctypes.data)
        if rc != 0 and need.value > tcap:
            buf = np.empty(need.value, np.uint8)
            rc = self.lib.nsc_tsv_stage(*args, buf.ctypes.data, need.value, C.byref(need), lens.ctypes.data, hashes.ctypes.data)
        _lib.check(rc)
        return buf[:need.value - 1].tobytes().decode().split("\n"), lens, hashes

    def decode_rows(self, rows):
        """uint8 [len(rows),5,32,23]: the matrices of the listed candidates only."""
        rows = np.ascontiguousarray(rows, np.int64)
        m = np.empty((len(rows), 5, 32, 23), np.uint8)
        if len(rows):
            _lib.check(self.lib.nsc_tsv_decode_rows(self.h, rows.ctypes.data, len(rows), m.ctypes.data, self.num_threads))
        return m

    def close(self):
        if getattr(self, "h", None):
            self.lib.nsc_tsv_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
