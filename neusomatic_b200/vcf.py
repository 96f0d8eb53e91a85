"""Dictionaries -> VCF: the host stage behind the scoring path (reference neusomatic/python/call.py:121-387 and
utils.py:84-95), restated on the in-memory pileup matrices this package keeps instead of PNG files.

Same entry points, argument meaning and results as the reference:

  * ``pred_vcf_records_path(record)``   call.py:121-305  one non-NONE candidate -> ``[path, [chrom, pos, ref, alt, prob,
                                        [path, pred]]]`` or ``[]``; ``record[1]`` may be the reference's PNG file name or the
                                        uint8 ``[5, 32, 3]`` matrix itself, ``record[5]`` a FASTA path or an open reader
  * ``pred_vcf_records``                call.py:308-333
  * ``pred_vcf_records_none``           call.py:336-354
  * ``get_vcf_records`` / ``write_vcf`` call.py:357-387
  * ``prob2phred``                      utils.py:84-95 (the arithmetic stays in the type of ``p``: np.float32 here)

Parity is pinned by ``tests/golden/vcf_*.npz`` (outputs of the UNMODIFIED reference functions and of the reference's whole
``call_neusomatic`` on synthetic candidates; ``tests/golden/make_golden_vcf.py``) and checked in ``tests/test_vcf.py``.
The expressions that decide a comparison are kept in the reference's form on purpose (Python ``sum`` / ``min`` / ``max`` /
``round`` over np.float32 scalars, comparisons against Python floats): the dictionaries hold np.float32 values and numpy's
scalar promotion decides borderline cases.
"""
from __future__ import annotations

import logging
import multiprocessing
import os
import pickle
import queue
import struct
import subprocess
import sys
import threading
import traceback

import numpy as np

from .fasta import fasta_index, open_fasta

ACGT = "ACGT"
# constants of the centre refinement (call.py:148-151)
MIN_ACCEPTABLE_PROBMAX = 0.05
CENTER_DIST_ROUNDBACK = 0.8
TOO_FAR_CENTER = 3
NEIGH_BASES = 7


def get_type(ref, alt):
    """call.py:43-51."""
    d = len(ref) - len(alt.split(",")[0])
    return "DEL" if d > 0 else ("INS" if d < 0 else "SNP")


def prob2phred(p, max_phred=100):
    """utils.py:84-95."""
    assert 0 <= p <= 1
    if p == 1:
        return max_phred
    if p == 0:
        return 0
    q = -10 * np.log10(1 - p)
    return max_phred if q > max_phred else q


def _load_image(src):
    """The ``non_transformed_matrices[i, :, :, 0:3]`` image of a candidate as float64 in [0, 1] (call.py:128)."""
    if isinstance(src, str):
        from imageio import imread            # the reference's PNG hand-over, if a caller still uses it
        src = imread(src)
    return np.asarray(src) / 255.0


class _Pileup:
    """Column bookkeeping of one candidate image: plane 0 = reference, plane 1 = tumor counts, rows = (gap, A, C, G, T)."""

    def __init__(self, image):
        self.I = image
        self.width = image.shape[1]
        ref = image[:, :, 0]
        top = np.argmax(ref, 0)
        depth = sum(ref, 0) > 0                # Python sum over the rows, as the reference
        self.gap_cols = np.where((top == 0) & depth)[0]       # zref_pos: reference gap = insertion columns
        self.base_cols = np.where((top > 0) & depth)[0]       # nzref_pos

    def ref_base(self, col):
        return ACGT[np.argmax(self.I[1:, col, 0])]

    def tumor_base(self, col):
        return ACGT[np.argmax(self.I[1:, col, 1])]

    def tumor_depth(self, col):
        return sum(self.I[1:, col, 1])

    def second_tumor_base(self, col):
        """The tumor base after the reference base's row is taken out (call.py:223-226)."""
        rb = np.argmax(self.I[1:, col, 0])
        t = self.I[1:, col, 1].copy()
        t[rb] = 0
        return ACGT[rb], ACGT[np.argmax(t)]


def _refine_center(pl, pred, center, ref, alt, vartype_classes):
    """call.py:152-191: walks the type probabilities from the largest down until one gives a centre column within
    TOO_FAR_CENTER of the candidate's.  Mutates ``pred[3]`` in place (a rejected class is zeroed), as the reference does on
    its shallow copy.  Returns (type or "NONE", centre column, ins_no_zref_pos, centre prediction)."""
    ins_no_gap = False
    center_pred = None
    while True:
        probs = pred[3]
        if sum(probs) < MIN_ACCEPTABLE_PROBMAX:
            break
        amx = np.argmax(probs)
        type_pred = vartype_classes[amx]
        if type_pred == "NONE":
            break
        center_pred = min(max(0, pred[1][0]), pl.width - 1)
        if abs(center_pred - center) < CENTER_DIST_ROUNDBACK:
            col = center
        else:
            col = int(round(center_pred))
        if type_pred != "INS" and col in pl.gap_cols:
            if center in pl.base_cols:
                col = center
            elif len(pl.base_cols) > 0:
                col = pl.base_cols[np.argmin(abs(pl.base_cols - center_pred))]
            else:
                break
        elif type_pred == "INS" and col in pl.base_cols:
            if len(pl.gap_cols) == 0:
                if len(alt) < len(ref) + NEIGH_BASES:
                    break
                col = center
                ins_no_gap = True
            elif center in pl.gap_cols:
                col = center
            else:
                col = pl.gap_cols[np.argmin(abs(pl.gap_cols - center_pred))]
        if abs(col - center) > TOO_FAR_CENTER:
            pred[3][amx] = 0
            continue
        pred[0] = type_pred
        return type_pred, col, ins_no_gap, center_pred
    return "NONE", None, ins_no_gap, center_pred


def _snp_alleles(pl, col, len_pred):
    """call.py:215-231."""
    ref_, alt_ = "", ""
    for i in range(len_pred):
        right = pl.base_cols[pl.base_cols >= (col + i)]
        if len(right) > 0:
            c = right[np.argmin(abs(right - (col + i)))]
            rb, ab = pl.second_tumor_base(c)
            ref_ += rb
            alt_ += ab
            if pl.tumor_depth(c) == 0:
                break
    return ref_, alt_


def pred_vcf_records_path(record):
    """call.py:121-305.  Returns the VCF record of one candidate predicted as a variant, ``[]`` when the prediction does
    not survive the checks, ``None`` after an exception (the reference logs and returns None; pred_vcf_records raises)."""
    path, true_path_, pred_all, chroms, vartype_classes, ref_file = record
    log = logging.getLogger("{} ({})".format(pred_vcf_records_path.__name__, multiprocessing.current_process().name))
    try:
        fasta = open_fasta(ref_file)
        pl = _Pileup(_load_image(true_path_))
        chrom, pos, ref, alt, _, center, _, _, _ = path.split(".")
        center, pos = int(center), int(pos)

        pred = list(pred_all)                  # shallow: pred[3] is the dictionary entry's own list
        type_pred, col, ins_no_gap, center_pred = _refine_center(pl, pred, center, ref, alt, vartype_classes)
        if type_pred == "NONE":
            return []
        len_pred = pred[2]

        # column <-> reference position, anchored at the candidate's own column (call.py:199-216)
        kind = get_type(ref, alt)
        anchor = {"DEL": (pos + 1, center), "INS": (pos, center - 1), "SNP": (pos, center)}[kind]
        col_2_pos = {c: j for j, c in enumerate(pl.base_cols)}
        if anchor[1] not in col_2_pos:
            return []
        shift = anchor[0] - col_2_pos[anchor[1]]
        for c in pl.base_cols:
            col_2_pos[c] += shift
        pos_2_col = {v: k for k, v in col_2_pos.items()}

        if not abs(center_pred - center) < TOO_FAR_CENTER:
            return []
        if type_pred == "SNP":
            pos_ = col_2_pos[col]
            ref_, alt_ = _snp_alleles(pl, col, len_pred)
            if not ref_:
                return []
        elif type_pred == "INS":
            if ins_no_gap:
                pos_, ref_, alt_ = pos, ref.upper(), alt.upper()
            else:
                pos_, left = -1, col - 1
                for left in range(col - 1, 0, -1):           # nearest reference column left of the insertion
                    if left in pl.base_cols:
                        pos_ = col_2_pos[left]
                        break
                if pos_ == -1:
                    return []
                if pl.tumor_depth(left) == 0:
                    return []
                ref_ = pl.ref_base(left)
                alt_ = ref_
                if len_pred == 3:                            # class 3 = "3 or longer": trust the candidate's length
                    len_pred = max(len(alt) - len(ref), len_pred)
                for c in range(left + 1, pl.width):
                    if c in pl.gap_cols:
                        alt_ += pl.tumor_base(c)
                    else:
                        break
                    if (len(alt_) - len(ref_)) >= len_pred:
                        break
        else:                                                # DEL
            pos_ = col_2_pos[col] - 1
            if pos_ not in pos_2_col:
                return []
            ref_ = pl.ref_base(pos_2_col[pos_])
            alt_ = ref_
            if len_pred == 3:
                len_pred = max(len(ref) - len(alt), len_pred)
            for c in range(col, pl.width):
                if c in pl.base_cols:
                    ref_ += pl.ref_base(c)
                if (len(ref_) - len(alt_)) >= len_pred:
                    break
            if (len(ref_) - len(alt_)) < len_pred:
                pos_, ref_, alt_ = pos, ref.upper(), alt.upper()
        chrom_ = chroms[int(chrom)]
        if fasta.fetch(chrom_, pos_ - 1, pos_ + len(ref_) - 1).upper() != ref_.upper():
            return []
        if ref_ == alt_:
            return []

        pred = list(pred_all)                  # the probabilities as the refinement loop left them
        if type_pred == "SNP":
            prob = pred[3][3] * (pred[4][1])
        elif type_pred == "INS":
            prob = pred[3][1] * (1 - pred[4][0]) if not ins_no_gap else pred[3][1]
        else:
            prob = pred[3][0] * (1 - pred[4][0])
        return [path, [chrom_, pos_, ref_, alt_, prob, [path, pred]]]
    except Exception as ex:      # noqa: BLE001  (reference behaviour: log, return None, the caller raises)
        log.error(traceback.format_exc())
        log.error(ex)
        return None


POOL_CHUNK = 512          # called variants per pool task
POOL_MIN = 2048           # the workers are started once this many records have been asked for


def _records_chunk(job):
    """Pool task: the records of one chunk of called variants.  The chunk arrives as arrays (tags, images, the columns of
    the table) -- pickling per-candidate lists of np.float32 scalars costs more than building the records."""
    paths, images, cols, chroms, vartype_classes, ref_file, index, with_pred = job
    from .predtable import PredTable
    fasta = open_fasta(ref_file, index)
    tab = PredTable(vartype_classes)
    tab.add(paths, *cols)
    out = []
    for i, path in enumerate(paths):
        o = pred_vcf_records_path([path, images[i], tab[path], chroms, vartype_classes, fasta])
        if o is None:
            return None
        if o and not with_pred:
            o[1][5] = None                 # [path, pred]: dropped by get_vcf_records, expensive to send back
        out.append(o)
    return out


def _worker_main():
    """A pool worker: length-prefixed pickled jobs on stdin, results on stdout, until stdin closes.  Started as a fresh
    interpreter (``python -c``): it imports numpy and this module, never torch or CUDA, and shares nothing with its parent."""
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr                # a stray print must not end up in the result stream
    while True:
        hdr = inp.read(8)
        if len(hdr) < 8:
            return
        job = pickle.loads(inp.read(struct.unpack("<Q", hdr)[0]))
        blob = pickle.dumps(_records_chunk(job), protocol=pickle.HIGHEST_PROTOCOL)
        out.write(struct.pack("<Q", len(blob)))
        out.write(blob)
        out.flush()


class _Workers:
    """``n`` worker interpreters fed from one queue (a feeder thread per worker: send a job, wait for its result)."""

    def __init__(self, n):
        self.jobs = queue.Queue()
        self.results, self.failed, self.cv = {}, None, threading.Condition()
        self.procs, self.threads = [], []
        # the interpreters are started from a thread of their own: the caller gets on with its model while they load
        self.launcher = threading.Thread(target=self._launch, args=(n,), daemon=True)
        self.launcher.start()

    def _launch(self, n):
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        cmd = [sys.executable, "-c", "from neusomatic_b200.vcf import _worker_main; _worker_main()"]
        try:
            for _ in range(n):
                p = subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, env=env)
                t = threading.Thread(target=self._feed, args=(p,), daemon=True)
                self.procs.append(p)
                self.threads.append(t)
                t.start()
        except Exception as ex:      # noqa: BLE001
            with self.cv:
                if not self.threads:                 # nobody to do the work: whoever waits must hear about it
                    self.failed = ex
                self.cv.notify_all()

    def _feed(self, proc):
        while True:
            item = self.jobs.get()
            if item is None:
                return
            seq, job = item
            try:
                blob = pickle.dumps(job, protocol=pickle.HIGHEST_PROTOCOL)
                proc.stdin.write(struct.pack("<Q", len(blob)))
                proc.stdin.write(blob)
                proc.stdin.flush()
                hdr = proc.stdout.read(8)
                if len(hdr) < 8:
                    raise RuntimeError("a VCF-stage worker exited (code %s)" % proc.poll())
                res = pickle.loads(proc.stdout.read(struct.unpack("<Q", hdr)[0]))
                with self.cv:
                    self.results[seq] = res
                    self.cv.notify_all()
            except Exception as ex:      # noqa: BLE001
                with self.cv:
                    self.failed = ex
                    self.cv.notify_all()
                return

    def wait(self, seqs):
        with self.cv:
            while self.failed is None and not all(q in self.results for q in seqs):
                self.cv.wait()
            if self.failed is not None:
                raise RuntimeError("VCF stage: %s" % self.failed)
            return [self.results.pop(q) for q in seqs]

    def stop(self, kill=False):
        self.launcher.join()
        for _ in self.threads:
            self.jobs.put(None)
        for p in self.procs:
            try:
                if kill:
                    p.kill()
                else:
                    p.stdin.close()
            except Exception:      # noqa: BLE001
                pass
        for p in self.procs:
            try:
                p.wait(timeout=10)
            except Exception:      # noqa: BLE001
                p.kill()
        for p in self.procs:
            for f in (p.stdin, p.stdout):
                try:
                    f.close()
                except Exception:      # noqa: BLE001
                    pass


class RecordPool:
    """call.py:308-333 as a service: ``submit(final_preds, true_path)`` starts building the VCF records of one group of
    candidates and returns a handle, ``collect(handle)`` returns them in the order of ``final_preds``.  With
    ``num_threads`` > 1, a PredTable and a FASTA path, the records are built by a pool of worker processes like the
    reference's -- fed with chunks of 512 candidates shipped as arrays, while the caller goes on scoring the next group.
    The workers are fresh interpreters (no fork of a process that holds a CUDA context and threads; they import numpy and
    this module only); they are started by ``prefork`` or once 2048 records have been asked for -- before that, and in every
    other case, ``submit`` builds the records itself."""

    def __init__(self, ref_file, chroms, vartype_classes, num_threads, with_pred=True):
        self.ref_file, self.chroms, self.classes = ref_file, chroms, vartype_classes
        self.num_threads, self.with_pred = int(num_threads), with_pred
        self.pool = None
        self.asked = 0
        self.seq = 0

    def prefork(self):
        """Start the workers now: they load while the caller sets up its model and scores the first candidates."""
        if self.pool is None and self.num_threads > 1 and isinstance(self.ref_file, str):
            self.pool = _Workers(min(self.num_threads, 32))

    def _serial(self, final_preds, true_path, paths):
        fasta = open_fasta(self.ref_file)
        out = []
        for path in paths:
            o = pred_vcf_records_path([path, true_path[path], final_preds[path], self.chroms, self.classes, fasta])
            if o is None:
                raise Exception("pred_vcf_records_path failed!")
            if o and not self.with_pred:
                o[1][5] = None
            out.append(o)
        return out

    def submit(self, final_preds, true_path):
        paths = list(final_preds.keys())
        self.asked += len(paths)
        poolable = self.num_threads > 1 and isinstance(self.ref_file, str) and hasattr(final_preds, "columns_of")
        if not poolable or (self.pool is None and self.asked < POOL_MIN) or not paths:
            return ("done", self._serial(final_preds, true_path, paths))
        self.prefork()
        index = fasta_index(open_fasta(self.ref_file))
        seqs = []
        for lo in range(0, len(paths), POOL_CHUNK):
            ps = paths[lo:lo + POOL_CHUNK]
            imgs = [true_path[p] for p in ps]
            if all(isinstance(im, np.ndarray) for im in imgs):
                imgs = np.stack(imgs)
            job = (ps, imgs, final_preds.columns_of(ps), self.chroms, self.classes, self.ref_file, index, self.with_pred)
            self.pool.jobs.put((self.seq, job))
            seqs.append(self.seq)
            self.seq += 1
        return ("pending", seqs)

    def collect(self, handle):
        kind, h = handle
        if kind == "pending":
            parts = self.pool.wait(h)
            if any(part is None for part in parts):
                raise Exception("pred_vcf_records_path failed!")
            h = [o for part in parts for o in part]
        return list(filter(None, h))

    def close(self, kill=False):
        if self.pool is not None:
            self.pool.stop(kill)
            self.pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close(kill=exc[0] is not None)
        return False


def pred_vcf_records(ref_file, final_preds, true_path, chroms, vartype_classes, num_threads, with_pred=True):
    """call.py:308-333.  ``true_path[path]`` holds the uint8 [5,32,3] matrix (or a PNG file name).  One-shot use of
    ``RecordPool`` (worker processes for 4 x 2048 records or more: starting them costs about as much as 3000 records).
    ``with_pred=False`` leaves out the trailing ``[path, pred]`` of every record, which ``get_vcf_records`` drops anyway
    (call.py:357-363)."""
    with RecordPool(ref_file, chroms, vartype_classes, num_threads, with_pred) as rp:
        rp.asked = POOL_MIN                    # one request: no later group to amortise the workers over, so decide now
        if len(final_preds) < 4 * POOL_MIN:
            rp.num_threads = 1
        return rp.collect(rp.submit(final_preds, true_path))


def pred_vcf_records_none(none_preds, chroms, with_pred=True):
    """call.py:336-354: candidates called NONE whose variant probability is still above 0.005 go to none.vcf.
    ``with_pred=False`` (PredTable only) leaves out the trailing ``[path, pred]`` that get_vcf_records drops."""
    if hasattr(none_preds, "none_records"):         # call.PredTable: the same filter on whole arrays
        return none_preds.none_records(chroms, with_pred)
    records = {}
    for path, pred in none_preds.items():
        chrom, pos, ref, alt, _, _, _, _, _ = path.split(".")
        if len(ref) == len(alt):
            prob = pred[3][3] * (pred[4][1])
        elif len(ref) < len(alt):
            prob = pred[3][1] * (1 - pred[4][0])
        else:
            prob = pred[3][0] * (1 - pred[4][0])
        if prob > 0.005:
            records[path] = [chroms[int(chrom)], pos, ref, alt, prob, [path, pred]]     # pos stays the tag's string
    return records.items()


def get_vcf_records(all_vcf_records):
    """call.py:357-363."""
    return [all_vcf_records[path][:-1] for path in all_vcf_records if all_vcf_records[path]]


def _vcf_lines_scalar(records, pass_threshold, lowqual_threshold):
    """call.py:366-387 record by record, in the reference's own expressions."""
    lines, seen = [], set()
    for chrom_, pos_, ref_, alt_, prob in records:
        if ref_ == alt_:
            continue
        filter_ = "PASS" if prob >= pass_threshold else ("LowQual" if prob >= lowqual_threshold else "REJECT")
        line = "\t".join([chrom_, str(pos_), ".", ref_, alt_, "{:.4f}".format(np.round(prob2phred(prob), 4)), filter_,
                          "SCORE={:.4f}".format(np.round(prob, 4)), "GT", "0/1"]) + "\n"
        if line not in seen:
            seen.add(line)
            lines.append(line)
    return lines


def vcf_lines(vcf_records, chroms_order, pass_threshold, lowqual_threshold):
    """The lines write_vcf emits (call.py:366-387), in order, first occurrence of a repeated line only.

    When every probability is an np.float32 scalar in [0, 1] (what this package's dictionaries hold), the two roundings and
    the threshold comparisons run on whole arrays: np.round on a float32 array is the same multiply / rint / divide as on a
    float32 scalar, and formatting goes through the same Python float.  log10 stays a scalar call per record -- an array
    loop may use a different (SIMD) implementation whose last bit can differ, which would show in the fourth decimal of QUAL.
    Anything else (Python floats, out-of-range values) takes the record-by-record path."""
    records = sorted((r for r in vcf_records if len(r) > 0), key=lambda x: [chroms_order[x[0]], x[1]])
    records = [r for r in records if r[2] != r[3]]
    if not records or not all(type(r[4]) is np.float32 for r in records):
        return _vcf_lines_scalar(records, pass_threshold, lowqual_threshold)
    prob = np.array([r[4] for r in records], np.float32)
    if not ((prob >= 0) & (prob <= 1)).all():
        return _vcf_lines_scalar(records, pass_threshold, lowqual_threshold)        # prob2phred asserts
    one_minus = np.float32(1) - prob
    log10 = np.log10
    q = np.empty(len(records), np.float32)
    for i, v in enumerate(one_minus):                      # utils.py:84-95 with p an np.float32: float32 throughout
        q[i] = -10 * log10(v) if v > 0 else 100
    q[prob == 0] = 0
    q = np.where(q > 100, np.float32(100), q)
    qual = np.round(q, 4).astype(np.float64).tolist()
    score = np.round(prob, 4).astype(np.float64).tolist()
    is_pass = (prob >= pass_threshold).tolist()
    is_low = (prob >= lowqual_threshold).tolist()
    lines, seen = [], set()
    for r, ql, sc, ps, lq in zip(records, qual, score, is_pass, is_low):
        line = "%s\t%s\t.\t%s\t%s\t%.4f\t%s\tSCORE=%.4f\tGT\t0/1\n" % (
            r[0], r[1], r[2], r[3], ql, "PASS" if ps else ("LowQual" if lq else "REJECT"), sc)
        if line not in seen:
            seen.add(line)
            lines.append(line)
    return lines


def write_vcf(vcf_records, output_vcf, chroms_order, pass_threshold, lowqual_threshold):
    """call.py:366-387."""
    with open(output_vcf, "w") as ov:
        ov.writelines(vcf_lines(vcf_records, chroms_order, pass_threshold, lowqual_threshold))
