"""``PredTable``: call.py's ``final_preds`` / ``none_preds`` dictionaries (reference neusomatic/python/call.py:96-114) held as
whole-batch arrays.  numpy only -- the pool workers of the VCF stage import it without pulling in torch."""
from __future__ import annotations

import bisect
import collections.abc

import numpy as np

VARTYPE_CLASSES = ["DEL", "INS", "NONE", "SNP"]          # call.py:407


class PredTable(collections.abc.Mapping):
    """``{path: [type, pos, len, probs_type, probs_len, logits_type, logits_len]}`` of call.py:96-114 held as whole-batch
    arrays.  Entries are materialised on access with the reference's element types (np.float32 scalars in the four lists,
    the [1]-shaped row of ``o2``, the numpy integer of the length argmax); a later candidate with the same tag replaces an
    earlier one and keeps its place in the iteration order, like assignment into a dict.  ``none_records`` is the
    whole-array form of pred_vcf_records_none (call.py:336-354)."""

    def __init__(self, vartype_classes=VARTYPE_CLASSES):
        self.classes = list(vartype_classes)
        self._paths = []                  # all rows, in arrival order
        self._parts = []                  # per add(): (a1, o2, a3, P1, P3, O1, O3, lens)
        self._cat = None
        self._index = {}
        self._indexed = 0
        self._hashes = []                 # per add(): uint64 hashes of the paths, or None
        self._starts = []                 # first row of every part
        self._kept = []                   # per add(): (rows with prob > 0.005, their prob) of the none.vcf filter, or None
        self._unique = None

    def add(self, paths, a1, o2, a3, p1, p3, o1, o3, lens=None, hashes=None):
        """Rows of one batch; the probability / logit arrays are rounded here (np.round on float32 == round(x, 4)).
        ``hashes``: 64-bit hashes of the paths (nsc_tsv_stage) -- if every row of the table came with one and they are
        pairwise distinct, no tag repeats and ``none_records`` need not build the tag dictionary to resolve duplicates."""
        if len(paths) == 0:
            return
        f32 = lambda a: np.asarray(a, np.float32)
        base = len(self._paths)
        self._paths.extend(paths)
        self._hashes.append(None if hashes is None else np.asarray(hashes, np.uint64))
        self._unique = None
        part = (np.asarray(a1, np.int64), f32(o2).reshape(len(paths), -1), np.asarray(a3, np.int64),
                np.round(f32(p1), 4), np.round(f32(p3), 4), np.round(f32(o1), 4), np.round(f32(o3), 4),
                None if lens is None else np.asarray(lens, np.int32))
        self._parts.append(part)
        self._starts.append(base)
        # the none.vcf filter of this batch (call.py:343-350), while its arrays are in cache
        if part[7] is not None:
            prob = self._none_prob(part[3], part[4], part[7])
            keep = np.nonzero(prob > 0.005)[0]
            self._kept.append((base + keep, prob[keep]))
        else:
            self._kept.append(None)
        self._cat = None

    @staticmethod
    def _none_prob(P1, P3, lens):
        """prob of call.py:343-350 by allele-length class, in float32 like the reference's np.float32 scalars."""
        rl, al = lens[:, 0], lens[:, 1]
        one = np.float32(1)
        return np.where(rl == al, P1[:, 3] * P3[:, 1], np.where(rl < al, P1[:, 1] * (one - P3[:, 0]), P1[:, 0] * (one - P3[:, 0])))

    def _arrays(self):
        if self._cat is None:
            parts = self._parts
            cols = [np.concatenate([p[j] for p in parts]) for j in range(7)] if parts else [np.zeros((0,))] * 7
            lens = None
            if parts and all(p[7] is not None for p in parts):
                lens = np.concatenate([p[7] for p in parts])
            self._cat = cols + [lens]
            self._parts = [tuple(self._cat)] if parts else []
            self._starts = [0] if parts else []
        return self._cat

    def _idx(self):
        n = len(self._paths)
        if self._indexed != n:
            self._index.update(zip(self._paths[self._indexed:], range(self._indexed, n)))
            self._indexed = n
        return self._index

    def _entry(self, r):
        # from the part that holds row r (no concatenation of the whole table for a handful of look-ups)
        k = bisect.bisect_right(self._starts, r) - 1
        a1, o2, a3, P1, P3, O1, O3, _ = self._parts[k]
        r = r - self._starts[k]
        return [self.classes[a1[r]], o2[r], a3[r], list(P1[r]), list(P3[r]), list(O1[r]), list(O3[r])]

    def __getitem__(self, path):
        return self._entry(self._idx()[path])

    def __iter__(self):
        return iter(self._idx())

    def __len__(self):
        return len(self._idx())

    def __contains__(self, path):
        return path in self._idx()

    def none_records(self, chroms, with_pred=True):
        """call.py:336-354 on arrays: prob by allele-length class in float32, keep prob > 0.005, materialise only those."""
        if not self._paths:
            return {}.items()
        if self._unique is None:
            self._unique = False
            if self._hashes and all(h is not None for h in self._hashes):
                h = np.sort(np.concatenate(self._hashes))
                self._unique = bool((h[1:] != h[:-1]).all())
        if self._unique and all(k is not None for k in self._kept):
            # no tag repeats: every row is its own entry, in arrival order, and the filter ran batch by batch in add()
            sel = np.concatenate([k[0] for k in self._kept])
            prob = np.concatenate([k[1] for k in self._kept])
        else:
            if self._unique:
                rows = np.arange(len(self._paths), dtype=np.int64)
            else:                             # a later row replaces an earlier one with the same tag, like assignment into a dict
                idx = self._idx()
                rows = np.fromiter(idx.values(), np.int64, len(idx))
            _, _, _, P1, P3, _, _, lens = self._arrays()
            if lens is None:
                lens = np.array([[len(f[2]), len(f[3])] for f in (p.split(".") for p in self._paths)], np.int32).reshape(-1, 2)
            prob = self._none_prob(P1[rows], P3[rows], lens[rows])
            keep = np.nonzero(prob > 0.005)[0]
            sel, prob = rows[keep], prob[keep]
        records = {}
        for r, pr in zip(sel, prob):
            path = self._paths[r]
            chrom, pos, ref, alt = path.split(".")[:4]
            records[path] = [chroms[int(chrom)], pos, ref, alt, pr, [path, self._entry(r)] if with_pred else None]
        return records.items()

    def columns_of(self, paths):
        """(a1, o2, a3, P1, P3, O1, O3) of the given tags' CURRENT rows (the later row of a repeated tag), as arrays: what
        ``add`` takes -- a chunk of the table can be rebuilt elsewhere (the pool workers of vcf.pred_vcf_records)."""
        idx = self._idx()
        rows = np.fromiter((idx[p] for p in paths), np.int64, len(paths))
        return tuple(c[rows] for c in self._arrays()[:7])

