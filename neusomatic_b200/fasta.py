"""Reference-genome access for the VCF stage: the three things the reference asks of ``pysam.FastaFile`` on this path
(call.py:125,276,403-405; utils.py:65-69) -- ``references``, ``fetch(chrom, start, end)`` (0-based, half open) and the
context-manager protocol.  ``open_fasta`` hands out pysam's reader when pysam is installed and this indexed plain-text
reader otherwise (a ``.fai`` next to the file is used if present, else the index is built by one scan)."""
from __future__ import annotations

import os


class FastaFile:
    def __init__(self, path, index=None):
        self.path = path
        self._index = {}            # name -> (length, offset of the first base, bases per line, bytes per line)
        fai = path + ".fai"
        if index is not None:       # handed over by a process that has already read or built it
            self._index = dict(index)
        elif os.path.exists(fai):
            for line in open(fai):
                f = line.rstrip("\n").split("\t")
                if len(f) >= 5:
                    self._index[f[0]] = (int(f[1]), int(f[2]), int(f[3]), int(f[4]))
        else:
            self._scan()
        self.references = tuple(self._index.keys())
        self.lengths = tuple(v[0] for v in self._index.values())
        self._f = open(path, "rb")

    def _scan(self):
        name, length, start, lb, lw, off = None, 0, 0, 0, 0, 0
        with open(self.path, "rb") as f:
            for raw in f:
                if raw.startswith(b">"):
                    if name is not None:
                        self._index[name] = (length, start, lb, lw)
                    name = raw[1:].split()[0].decode()
                    length, start, lb, lw = 0, off + len(raw), 0, 0
                else:
                    n = len(raw.rstrip(b"\r\n"))
                    if lb == 0 and n:
                        lb, lw = n, len(raw)
                    length += n
                off += len(raw)
        if name is not None:
            self._index[name] = (length, start, lb, lw)

    def fetch(self, reference, start=0, end=None):
        length, first, lb, lw = self._index[reference]
        start = max(int(start), 0)
        end = length if end is None else min(int(end), length)
        if end <= start or lb == 0:
            return ""
        b0 = first + (start // lb) * lw + start % lb
        b1 = first + ((end - 1) // lb) * lw + (end - 1) % lb + 1
        self._f.seek(b0)
        return self._f.read(b1 - b0).replace(b"\n", b"").replace(b"\r", b"").decode()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class MemoryFasta:
    """The same interface over sequences held in memory ({name: sequence}); used by the tests and the benchmark."""

    def __init__(self, seqs):
        self.seqs = dict(seqs)
        self.references = tuple(self.seqs.keys())
        self.lengths = tuple(len(s) for s in self.seqs.values())

    def fetch(self, reference, start=0, end=None):
        return self.seqs[reference][max(int(start), 0):end]

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_OPEN = {}


def open_fasta(ref, index=None):
    """``ref``: a path (one reader per path and process is kept open) or an object that already has ``fetch``.
    ``index``: the ``fasta_index`` of another process's reader of the same file (saves the scan when there is no .fai)."""
    if hasattr(ref, "fetch"):
        return ref
    key = (os.getpid(), ref)
    r = _OPEN.get(key)
    if r is None:
        try:
            import pysam
            r = pysam.FastaFile(ref)
        except ImportError:
            r = FastaFile(ref, index)
        _OPEN[key] = r
    return r


def fasta_index(reader):
    """The index of this module's own reader (None for pysam's, which reads the .fai itself)."""
    return getattr(reader, "_index", None)
