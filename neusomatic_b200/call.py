"""Host-side mirror of the scoring part of the reference's ``call.py``.

  * ``call_variants``       -- call.py:54-118, same signature and the same
                               ``(final_preds, none_preds, true_path)`` dictionaries, but one
                               fused device call + one D2H per batch and a vectorised
                               post-loop instead of 2 kernels + 4 syncs per candidate.
  * ``load_checkpoint``     -- call.py:424-460 (``module.`` prefix reconciliation, coverage_thr,
                               normalize_channels).
  * ``CandidateScorer``     -- the compact path: uint8 matrices + tag fields in host memory ->
                               ``nsc_score_host_u8`` (channel assembly fused on the device).
  * ``shard_ranges``        -- contiguous per-rank candidate ranges for multi-GPU scoring
                               (no collective; replaces nn.DataParallel of call.py:417-419).
  * ``PredTable``           -- the same ``{path: entry}`` dictionaries kept as whole-batch arrays and materialised
                               per entry on access (the reference builds ~40 Python objects per candidate; most
                               candidates are NONE and are only ever looked at by the none.vcf filter).
  * ``call_neusomatic``     -- call.py:389-554, same signature: candidate TSVs + FASTA + checkpoint -> ``pred.vcf`` /
                               ``none.vcf``, decode / score / post-processing overlapped.
  * ``pred_vcf_records*``, ``get_vcf_records``, ``write_vcf``, ``get_type`` -- re-exported from ``vcf.py``
                               (call.py:121-387).
"""
from __future__ import annotations

import concurrent.futures
import gc
import logging
import os
import time

import numpy as np
import torch

from .engine import Engine, VARTYPE_CLASSES, anns_to_255
from .predtable import PredTable                                                             # noqa: F401  (re-exported)
from .fasta import open_fasta
from .vcf import (RecordPool, get_type, get_vcf_records, pred_vcf_records, pred_vcf_records_none,      # noqa: F401  (reference names)
                  pred_vcf_records_path, vcf_lines, write_vcf)

logger = logging.getLogger(__name__)


def load_checkpoint(path_or_dict, map_location="cpu"):
    """Returns (state_dict without ``module.`` prefix, meta dict) -- call.py:424-460."""
    ck = path_or_dict
    if not isinstance(ck, dict):
        ck = torch.load(path_or_dict, map_location=map_location, weights_only=True)
    sd = {}
    for k, v in ck["state_dict"].items():
        sd[k[len("module."):] if k.startswith("module.") else k] = v
    meta = {
        "tag": ck.get("tag", ""),
        "epoch": ck.get("epoch", 0),
        "coverage_thr": ck["coverage_thr"],                                # call.py:434-436
        "normalize_channels": bool(ck.get("normalize_channels", False)),   # call.py:437-439
    }
    return sd, meta


def build_pred_dicts(o1, o2, o3, p1, p3, arg1, arg3, paths, vartype_classes=VARTYPE_CLASSES,
                     final_preds=None, none_preds=None):
    """Vectorised restatement of call.py:86-114 on host arrays.

    ``round(np.float32, 4)`` of the reference is numpy's float32 rounding; ``np.round`` on a
    float32 array applies the same ufunc element-wise."""
    final_preds = {} if final_preds is None else final_preds
    none_preds = {} if none_preds is None else none_preds
    o1 = np.asarray(o1, np.float32)
    o2 = np.asarray(o2, np.float32)
    o3 = np.asarray(o3, np.float32)
    r_p1, r_p3 = np.round(np.asarray(p1, np.float32), 4), np.round(np.asarray(p3, np.float32), 4)
    r_o1, r_o3 = np.round(o1, 4), np.round(o3, 4)
    # The entries keep the reference's element types (np.float32 scalars in the four lists, the [1]-shaped row of o2,
    # the np integer of argmax) but are cut out of whole-array conversions: iterating a flat array creates the scalars
    # in one pass, and a list slice replaces four per-candidate row views + list() calls (15 -> ~2 us per candidate).
    P1, P3, O1, O3 = list(r_p1.ravel()), list(r_p3.ravel()), list(r_o1.ravel()), list(r_o3.ravel())
    O2 = list(o2.reshape(len(o2), -1))
    A3 = list(np.asarray(arg3))
    A1 = np.asarray(arg1).tolist()
    none_idx = vartype_classes.index("NONE") if "NONE" in vartype_classes else -1
    # millions of small lists: the cyclic collector would re-scan the growing heap on every generation threshold
    # (4.5 of the 9 us per candidate); nothing allocated here can form a cycle, so it is paused for the loop
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        for i, path_ in enumerate(paths):
            path = path_[path_.rfind("/") + 1:]
            a = A1[i]
            j = 4 * i
            entry = [vartype_classes[a], O2[i], A3[i], P1[j:j + 4], P3[j:j + 4], O1[j:j + 4], O3[j:j + 4]]
            if a != none_idx:
                final_preds[path] = entry
            else:
                none_preds[path] = entry
    finally:
        if gc_was_on:
            gc.enable()
    return final_preds, none_preds


def non_transformed_images(matrices, center):
    """dataloader.py:398-404: uint8 [n,5,32,3] = (reference plane, tumor plane, centre marker) of compact candidates --
    the image call.py:89-95 hands to pred_vcf_records_path (there through a PNG file)."""
    matrices = np.asarray(matrices)
    n = matrices.shape[0]
    out = np.zeros((n, 5, 32, 3), np.uint8)
    out[..., 0:2] = matrices[..., 0:2]
    mx = matrices[..., 0].reshape(n, -1).max(1) if n else np.zeros(0, np.uint8)
    out[np.arange(n), :, np.asarray(center, np.int64), 2] = mx[:, None]
    return out


def _score_batch(net, matrices):
    """One batch through the drop-in module -> host array [B,19] = (o1 4, o2 1, o3 4, softmax(o1) 4, softmax(o3) 4, argmax
    o1, argmax o3): logits, probabilities and argmax all come out of the library's fused head kernel, one D2H per batch.
    Under nn.DataParallel (call.py:417-419) the batch is split over the module's devices by hand: every device runs its
    own engine on its slice (launches are asynchronous, so the devices overlap)."""
    inner, devices = net, None
    if isinstance(net, torch.nn.DataParallel):
        inner, devices = net.module, list(net.device_ids)
    if not hasattr(inner, "score"):
        raise TypeError("call_variants needs a neusomatic_b200.network.NeuSomaticNet")
    if not devices or len(devices) == 1 or matrices.shape[0] < 2 * len(devices):
        dev = next(inner.parameters()).device
        outs = [inner.score(matrices.to(dev, non_blocking=True))]
    else:
        outs = [inner.score(c.to("cuda:%d" % d, non_blocking=True))
                for c, d in zip(matrices.chunk(len(devices)), devices)]
    parts = []
    for o1, o2, o3, p1, p3, a1, a3 in outs:
        parts.append(torch.cat([o1, o2, o3, p1, p3, a1.float()[:, None], a3.float()[:, None]], 1).cpu())
    return torch.cat(parts, 0).numpy()


# candidates gathered from successive loader batches before one device call: the reference's batch sizes (100 in
# test/run_test.sh, 1000 by default, call.py:585-586) are one wave of 8-candidate groups or less on 148 SMs and run at the
# latency of ten dependent layer kernels; the results do not depend on how candidates are batched, so the loop scores
# them in larger slices
COALESCE_CANDIDATES = 8192


def call_variants(net, vartype_classes, call_loader, out_dir, model_tag, use_cuda):
    """call.py:54-118, same signature and return value ``(final_preds, none_preds, true_path)``.

    ``net`` is a ``neusomatic_b200.network.NeuSomaticNet`` (possibly wrapped in nn.DataParallel).  Loader batches are
    copied to the device as they arrive and scored in slices of at least COALESCE_CANDIDATES candidates (one fused library
    call and one D2H per slice); the dictionaries are filled in loader order, so a repeated tag is overwritten exactly as in
    the reference.  ``final_preds`` / ``none_preds`` are PredTables: read-only mappings whose entries are materialised on
    access with the reference's format and element types -- everything the reference does with them downstream
    (``.keys()``, ``[path]``: call.py:312-314, 340-343) works, while the ~40 Python objects per candidate that bound the loop
    at 0.3-0.4 M candidates/s are only built for the entries somebody reads; ``dict(final_preds)`` gives plain dicts.  ``true_path`` maps every non-NONE candidate to its ``non_transformed[:, :, 0:3]`` image.  When the
    reference's hand-over directory ``{out_dir}/matrices_{model_tag}`` exists and ``imageio`` is installed -- i.e. inside
    the reference's own call_neusomatic -- the image is written there as a PNG and ``true_path`` holds the file name,
    exactly as call.py:89-95 does, so the unmodified ``pred_vcf_records_path`` (call.py:128 ``imread``) keeps working.
    Otherwise ``true_path`` holds the uint8 [5,32,3] array itself, which this package's ``pred_vcf_records_path`` accepts
    (the PNG is a lossless round trip of exactly these bytes)."""
    if not use_cuda:
        raise RuntimeError("neusomatic_b200 scores on CUDA only")
    net.eval()
    final_preds, none_preds, true_path = PredTable(vartype_classes), PredTable(vartype_classes), {}
    none_idx = vartype_classes.index("NONE") if "NONE" in vartype_classes else -1
    png_dir, imwrite = None, None
    if out_dir is not None:
        d = "{}/matrices_{}".format(out_dir, model_tag)
        if os.path.isdir(d):
            try:
                from imageio import imwrite
                png_dir = d
            except ImportError:
                png_dir = None
    inner = net.module if isinstance(net, torch.nn.DataParallel) else net
    dev = next(inner.parameters()).device
    pending, n_pending = [], 0          # (device matrices, paths, non_transformed) of the loader batches not yet scored

    def flush():
        nonlocal pending, n_pending
        if not pending:
            return
        mats = pending[0][0] if len(pending) == 1 else torch.cat([p[0] for p in pending], 0)
        with torch.no_grad():
            packed = _score_batch(net, mats)
        lo = 0
        for _m, paths, nt in pending:
            hi = lo + len(paths)
            pk = packed[lo:hi]
            lo = hi
            if any("/" in p for p in paths):
                paths = [p.split("/")[-1] for p in paths]
            a1 = pk[:, 17].astype(np.int64)
            var = np.nonzero(a1 != none_idx)[0]
            non = np.nonzero(a1 == none_idx)[0]
            for tab, idx in ((final_preds, var), (none_preds, non)):
                if len(idx):
                    tab.add([paths[i] for i in idx], a1[idx], pk[idx, 4:5], pk[idx, 18].astype(np.int64), pk[idx, 9:13], pk[idx, 13:17],
                            pk[idx, 0:4], pk[idx, 5:9])
            if len(var) == 0:
                continue
            nt = np.asarray(nt)
            for i in var:
                path = paths[i]
                if png_dir is not None:
                    file_name = "{}/{}.png".format(png_dir, path)
                    if not os.path.exists(file_name):
                        imwrite(file_name, np.array(nt[i, :, :, 0:3]))
                    true_path[path] = file_name
                elif path not in true_path:
                    true_path[path] = np.array(nt[i, :, :, 0:3])
        pending, n_pending = [], 0

    iii = j = 0
    for data in call_loader:
        (matrices, labels, var_pos_s, var_len_s, non_transformed_matrices), (paths) = data
        iii += 1
        j += len(paths[0])
        pending.append((matrices.to(dev, non_blocking=True), list(paths[0]), non_transformed_matrices))
        n_pending += len(paths[0])
        if n_pending >= COALESCE_CANDIDATES:
            flush()
        if iii % 10 == 0:
            logger.info("Called {} candidates in this batch.".format(j))
    flush()
    logger.info("Called {} candidates in this batch.".format(j))
    return final_preds, none_preds, true_path


def shard_ranges(n, world_size):
    """Contiguous candidate ranges per rank: sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def call_sharded(cand, score_fn, rank=None, world_size=None, group=None):
    """Multi-GPU replacement of nn.DataParallel in call.py:417-419: every rank scores a contiguous range of
    the candidate stream (``shard_ranges``) with ``score_fn(Candidates) -> (final_preds, none_preds)`` and the
    dictionaries are merged on rank 0 in candidate order (later duplicates of a tag win, as in call.py:96-114).
    There is no collective on the data path; the only exchange is the host-side gather of the results.
    Returns (final_preds, none_preds) on rank 0 and (None, None) elsewhere."""
    import torch.distributed as dist
    from .synth import Candidates
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside [0, %d)" % (rank, world_size))
    lo, hi = shard_ranges(len(cand), world_size)[rank]
    part = Candidates(cand.matrices[lo:hi], cand.center[lo:hi], cand.tumor_cov[lo:hi], cand.normal_cov[lo:hi],
                      None if cand.anns is None else cand.anns[lo:hi], cand.tags[lo:hi])
    if getattr(cand, "anns255", None) is not None:
        part.anns255 = cand.anns255[lo:hi]
    mine = score_fn(part)
    if world_size == 1:
        return mine
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None, None
    final_preds, none_preds = {}, {}
    for f, n in gathered:                    # rank order == candidate order
        for k, v in f.items():
            none_preds.pop(k, None)
            final_preds[k] = v
        for k, v in n.items():
            final_preds.pop(k, None)
            none_preds[k] = v
    return final_preds, none_preds


class CandidateScorer:
    """Scores compact candidates (uint8 [N,5,32,23] + center/coverages [+ annotations]) held in
    host memory through ``nsc_score_host_u8``: the B200 replacement of the
    DataLoader -> ``.cuda()`` -> ``net()`` -> per-candidate softmax chain of call.py:68-114."""

    def __init__(self, checkpoint, ensemble=None, device=0, precision=None):
        sd, meta = load_checkpoint(checkpoint)
        C = sd["conv1.weight"].shape[1]
        if ensemble is not None and C != (119 if ensemble else 26):
            raise ValueError("checkpoint has %d input channels" % C)
        self.meta = meta
        self.coverage_thr = float(meta["coverage_thr"])
        if precision is None and os.environ.get("NSC_PRECISION"):      # NSC_PREC_* for entry points without a precision argument
            precision = int(os.environ["NSC_PRECISION"])
        self.engine = Engine(sd, C, device=device, precision=precision,
                             normalize_channels=meta["normalize_channels"])
        self.C = C

    def score(self, matrices, center, tumor_cov, normal_cov, anns=None, anns255=None):
        """Returns dict of numpy arrays: o1,o2,o3 (logits), p1,p3 (softmax), arg1,arg3.
        ``anns`` are the annotation values of the TSV (dataloader.py:44-58); ``anns255`` the pre-scaled
        float(double(ann) * 255.0) the C++ decoder emits."""
        n = matrices.shape[0]
        m8 = np.ascontiguousarray(matrices, np.uint8)
        meta = np.ascontiguousarray(np.stack([center, tumor_cov, normal_cov], 1), np.int32)
        a255 = None
        if self.C > 26:
            if anns is None and anns255 is None:
                raise ValueError("ensemble model needs annotations")
            a255 = np.ascontiguousarray(anns255 if anns255 is not None else anns_to_255(anns), np.float32)
        outs = self.engine.score_host_u8(m8, meta, a255, self.coverage_thr, extras=True)
        names = ["o1", "o2", "o3", "p1", "p3", "arg1", "arg3"]
        res = {k: t.numpy() for k, t in zip(names, outs)}
        assert res["o1"].shape[0] == n
        return res

    def call(self, cand, final_preds=None, none_preds=None):
        """Candidates (neusomatic_b200.synth.Candidates-like) -> (final_preds, none_preds)."""
        r = self.score(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns,
                       getattr(cand, "anns255", None))
        return build_pred_dicts(r["o1"], r["o2"], r["o3"], r["p1"], r["p3"], r["arg1"], r["arg3"], cand.tags,
                                VARTYPE_CLASSES, final_preds, none_preds)

    def call_table(self, cand, final_preds=None, none_preds=None, true_path=None):
        """Like ``call`` but into PredTables (arrays, entries on access); fills ``true_path`` with the images of the
        non-NONE candidates when a dict is given."""
        final_preds = PredTable() if final_preds is None else final_preds
        none_preds = PredTable() if none_preds is None else none_preds
        r = self.score(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns, getattr(cand, "anns255", None))
        _split_into_tables(r, cand.tags, None, cand.matrices, cand.center, final_preds, none_preds, true_path)
        return final_preds, none_preds

    def call_tsv(self, path, batch=65536, num_threads=0, final_preds=None, none_preds=None, true_path=None):
        """call.py:504-534 for one candidates_*.tsv: C++ decode (csrc/nsc_tsv.cpp) -> nsc_score_host_u8 -> PredTables
        ``(final_preds, none_preds)``, in slices of ``batch`` candidates, the three stages overlapped (the decoder threads
        fill slot k+1 while the GPU scores slot k and the host files slot k-1)."""
        final_preds = PredTable() if final_preds is None else final_preds
        none_preds = PredTable() if none_preds is None else none_preds
        run_tsv_pipeline([self], [(path, 0, None)], batch, num_threads, final_preds, none_preds, true_path)
        return final_preds, none_preds


def _split_into_tables(r, tags, lens, matrices, center, final_preds, none_preds, true_path, hashes=None):
    """Files one scored slice: rows predicted NONE go to ``none_preds``, the others to ``final_preds`` (call.py:88-114);
    ``true_path`` receives the non_transformed image of every non-NONE candidate (call.py:89-95)."""
    a1 = np.asarray(r["arg1"])
    none_idx = VARTYPE_CLASSES.index("NONE")
    var = np.nonzero(a1 != none_idx)[0]
    if any("/" in t for t in tags[:1]) or (len(tags) and "/" in tags[-1]):
        tags = [t[t.rfind("/") + 1:] for t in tags]                  # path_.split("/")[-1]
    cols = ("arg1", "o2", "arg3", "p1", "p3", "o1", "o3")
    if len(var) == 0:
        none_preds.add(tags, *[np.array(r[c]) for c in cols], lens=lens, hashes=hashes)
        return
    vt = [tags[i] for i in var]
    final_preds.add(vt, *[np.asarray(r[c])[var] for c in cols], lens=None if lens is None else lens[var])
    if 8 * len(var) < len(tags):             # few called variants (the usual case): cut them out instead of gathering the rest
        if len(var) <= 64:
            nt = list(tags)
            for i in var[::-1]:                  # (each del moves the tail of the list: only worth it for a handful)
                del nt[i]
        else:
            keep = np.ones(len(tags), bool)
            keep[var] = False
            nt = [t for t, k in zip(tags, keep.tolist()) if k]
        cut = lambda a: None if a is None else np.delete(np.asarray(a), var, axis=0)
        none_preds.add(nt, *[cut(r[c]) for c in cols], lens=cut(lens), hashes=cut(hashes))
    else:
        non = np.nonzero(a1 == none_idx)[0]
        sel = lambda a: None if a is None else np.asarray(a)[non]
        none_preds.add([tags[i] for i in non], *[sel(r[c]) for c in cols], lens=sel(lens), hashes=sel(hashes))
    if true_path is not None:
        # ``matrices`` may be a function of the row indices (the device-inflate pipeline keeps no decoded matrices on the host:
        # only the called variants' images are decoded here)
        mv = matrices(var) if callable(matrices) else np.asarray(matrices)[var]
        img = non_transformed_images(mv, np.asarray(center)[var])
        for t, im in zip(vt, img):
            true_path[t] = im


class _Slot:
    """Pinned host buffers of one in-flight slice: decoder output = scorer input, and the scorer's host results."""

    def __init__(self, n, num_anns, text_bytes=0):
        pin = torch.cuda.is_available()
        mk = lambda shape, dt: (torch.empty(shape, dtype=dt).pin_memory() if pin else torch.empty(shape, dtype=dt))
        self.n = n
        self.text_bytes = text_bytes
        self.tags = None              # (tag strings, allele lengths, tag hashes) of the slice staged in this slot
        if text_bytes:        # device inflate: the slice's lines as they are in the file + (offset, characters) of each image field
            self.text = mk((text_bytes,), torch.uint8)
            self.spans = mk((n, 2), torch.int32)         # read as uint32 by the library
            self.matrices = None
        else:
            self.matrices = mk((n, 5, 32, 23), torch.uint8)
        self.meta = mk((n, 3), torch.int32)
        self.anns = mk((n, num_anns), torch.float32) if num_anns else None
        self.out = Engine._host_out(n, True, pin)


def run_tsv_pipeline(scorers, pieces, batch, num_threads, final_preds, none_preds, true_path=None):
    """Decode -> score -> file, overlapped.  ``pieces`` = [(tsv path, first candidate, count or None)]; slices of at most
    ``batch`` candidates go round-robin to ``scorers`` (one per GPU: the path shards by candidate, no collective).  One
    decoder thread (the C++ decoder fans out over ``num_threads`` itself), one scoring thread per GPU and one filing
    thread; the ctypes calls release the GIL, so only the filing stage competes for it.

    By default the image fields are inflated ON THE DEVICE (``nsc_tsv_stage`` -> ``nsc_score_host_b64``,
    csrc/nsc_inflate_gpu.cu): the host stage only splits lines and parses tags.  ``NSC_TSV_HOST_DECODE=1`` keeps the
    inflate on the host threads (``nsc_tsv_decode`` -> ``nsc_score_host_u8``); the results are the same bytes."""
    from .tsv import CandidateTSV
    num_anns = scorers[0].C - 26
    device_inflate = os.environ.get("NSC_TSV_HOST_DECODE", "0") in ("", "0")
    tsvs, slices = {}, []
    times = {"open": 0.0, "decode": 0.0, "tags": 0.0, "score": 0.0, "file": 0.0, "total": 0.0}      # busy seconds per stage (PIPELINE_TIMES)
    t_begin = time.perf_counter()

    try:
        for path, first, count in pieces:
            if path not in tsvs:
                t0 = time.perf_counter()
                tsvs[path] = CandidateTSV(path, num_anns=num_anns, num_threads=num_threads)
                times["open"] += time.perf_counter() - t0
            n = len(tsvs[path]) - first if count is None else count
            for f0 in range(first, first + n, batch):
                slices.append((path, f0, min(batch, first + n - f0)))
        if not slices:
            return
        cap = max(s[2] for s in slices)
        n_slots = min(len(slices), 2 + len(scorers))
        text_cap = 0
        if device_inflate:
            text_cap = max(tsvs[p].range_bytes(f0, n) for p, f0, n in slices) + 64
        # pinned slot buffers are kept on the first scorer between calls (cudaHostAlloc of a few hundred MB is slow)
        cache = getattr(scorers[0], "_slot_cache", None)
        if (cache is None or cache[0] < cap or len(cache[1]) < n_slots or cache[2] != num_anns or
                (cache[1][0].text_bytes < text_cap) or (cache[1][0].text_bytes > 0) != device_inflate):
            text_cap += text_cap // 8
            cache = (cap, [_Slot(cap, num_anns, text_cap) for _ in range(n_slots)], num_anns)
            try:
                scorers[0]._slot_cache = cache
            except AttributeError:
                pass
        slots = cache[1][:n_slots]

        def decode(k, wait_for):
            if wait_for is not None:
                wait_for.result()                       # the slot's previous slice has been filed
            path, f0, n = slices[k]
            slot = slots[k % len(slots)]
            t = tsvs[path]
            t0 = time.perf_counter()
            try:
                if device_inflate:        # the staging threads gather the tag strings and allele lengths in the same pass
                    slot.tags = t.stage(f0, n, slot.text, slot.spans[:n], slot.meta[:n], None if slot.anns is None else slot.anns[:n],
                                        with_tags=True)
                    return t.range_bytes(f0, n)
                t.decode(f0, n, out={"matrices": slot.matrices[:n].numpy(), "meta": slot.meta[:n].numpy(),
                                     "anns255": None if slot.anns is None else slot.anns[:n].numpy()}, with_tags=False)
                return None
            finally:
                times["decode"] += time.perf_counter() - t0

        def score(k, fdec):
            nbytes = fdec.result()
            n = slices[k][2]
            slot = slots[k % len(slots)]
            sc = scorers[k % len(scorers)]
            outs = [None if t is None else t[:n] for t in slot.out]
            t0 = time.perf_counter()
            if device_inflate:
                sc.engine.score_host_b64(slot.text[:nbytes], slot.spans[:n], slot.meta[:n], None if slot.anns is None else slot.anns[:n],
                                         sc.coverage_thr, extras=True, out=outs)
            else:
                sc.engine.score_host_u8(slot.matrices[:n], slot.meta[:n], None if slot.anns is None else slot.anns[:n],
                                        sc.coverage_thr, extras=True, out=outs)
            times["score"] += time.perf_counter() - t0
            return outs

        def read_tags(k, wait_for):
            # the tag strings (the dictionaries' keys) and allele lengths of a slice, on a thread of their own: one C pass over
            # the lines (GIL released) and one split
            if device_inflate:
                return None                      # came with the staging pass (slot.tags)
            if wait_for is not None:
                wait_for.result()
            path, f0, n = slices[k]
            t0 = time.perf_counter()
            try:
                return tsvs[path].tags_lens(f0, n) + (None,)
            finally:
                times["tags"] += time.perf_counter() - t0

        def file_slice(k, fsc, ftags):
            path, f0, n = slices[k]
            tl = ftags.result()
            outs = fsc.result()                  # (after the slice's decode stage: slot.tags is this slice's)
            tags, lens, hashes = tl if tl is not None else slots[k % len(slots)].tags
            t0 = time.perf_counter()
            slot = slots[k % len(slots)]
            names = ["o1", "o2", "o3", "p1", "p3", "arg1", "arg3"]
            r = {nm: t.numpy() for nm, t in zip(names, outs)}
            mats = (lambda rows: tsvs[path].decode_rows(f0 + np.asarray(rows, np.int64))) if device_inflate else slot.matrices[:n].numpy()
            _split_into_tables(r, tags, lens, mats, slot.meta[:n, 0].numpy(), final_preds, none_preds, true_path, hashes)
            times["file"] += time.perf_counter() - t0

        TPE = concurrent.futures.ThreadPoolExecutor
        with TPE(1) as dec_ex, TPE(1) as file_ex, TPE(1) as tag_ex:
            sc_ex = [TPE(1) for _ in scorers]
            try:
                filed = []
                for k in range(len(slices)):
                    wait_for = filed[k - len(slots)] if k >= len(slots) else None
                    fd = dec_ex.submit(decode, k, wait_for)
                    fs = sc_ex[k % len(scorers)].submit(score, k, fd)
                    ft = tag_ex.submit(read_tags, k, wait_for)
                    filed.append(file_ex.submit(file_slice, k, fs, ft))
                for f in filed:
                    f.result()
            finally:
                for ex in sc_ex:
                    ex.shutdown(wait=True)
    finally:
        for t in tsvs.values():
            t.close()
        times["total"] = time.perf_counter() - t_begin
        PIPELINE_TIMES.clear()
        PIPELINE_TIMES.update(times)


PIPELINE_TIMES = {}      # busy seconds per stage of the last run_tsv_pipeline call (waits for the neighbouring stages excluded)


PREFORK_CANDIDATES = 200000       # calls at least this large start the VCF stage's process pool up front
RUN_CANDIDATES = 1000000          # consecutive groups share one run of the pipeline up to this many candidates


def plan_call_groups(counts, max_load_candidates):
    """The grouping of call.py:468-507 on candidate counts: files longer than max_load_candidates / 2 are cut into pieces
    of max(1, max_load_candidates / 2) candidates (the reference re-writes them with merge_tsvs; here a piece is a range of
    the same file), the pieces are sorted by size and taken in groups that end as soon as they hold more than
    max_load_candidates / 10 candidates.  ``counts`` = [(path, candidates)]; returns [[(path, first, count), ...], ...]."""
    pieces = []
    for path, L in counts:
        if L + 1 > max_load_candidates / 2:
            per = int(max(1, max_load_candidates / 2))
            pieces += [(path, f0, min(per, L - f0)) for f0 in range(0, L, per)]
        else:
            pieces.append((path, 0, L))
    pieces.sort(key=lambda x: x[2])
    groups, cur, cur_l = [], [], 0
    for i, pc in enumerate(pieces):
        cur.append(pc)
        cur_l += pc[2]
        if cur_l > max_load_candidates / 10 or i == len(pieces) - 1:
            groups.append(cur)
            cur, cur_l = [], 0
    return groups


def call_neusomatic(candidates_tsv, ref_file, out_dir, checkpoint, num_threads, batch_size, max_load_candidates,
                    pass_threshold, lowqual_threshold, ensemble, use_cuda):
    """call.py:389-554, same arguments and outputs (``{out_dir}/pred.vcf`` -- returned -- and ``{out_dir}/none.vcf``).

    Differences in mechanism, none in results: the candidates' lines are staged by the C++ reader into pinned batches, their
    image fields inflated on the device and scored (``nsc_score_host_b64``) on every visible GPU (candidate slices
    round-robin; the reference wraps the module in nn.DataParallel, call.py:417-419); oversized TSVs are read in ranges
    instead of being re-written by merge_tsvs, and consecutive groups of call.py:468-507 share one run of the pipeline; the
    pileup images go to pred_vcf_records_path in memory instead of through PNG files, and the records are built by the
    process pool while later candidates are being scored; ``batch_size`` only bounds the slice size from below (a slice is
    the unit of the stage / score / file pipeline, not of the arithmetic)."""
    from .tsv import CandidateTSV
    if not use_cuda:
        raise RuntimeError("neusomatic_b200 scores on CUDA only (no CPU fallback); the reference's CPU path is call.py itself")
    logger.info("-----------------Call Somatic Mutations--------------------")
    t_start = time.perf_counter()
    fasta = open_fasta(ref_file)
    chroms = list(fasta.references)
    chroms_order = {c: i for i, c in enumerate(chroms)}
    vartype_classes = list(VARTYPE_CLASSES)
    ck = checkpoint if isinstance(checkpoint, dict) else torch.load(checkpoint, map_location="cpu", weights_only=True)
    num_channels = load_checkpoint(ck)[0]["conv1.weight"].shape[1]
    counts = []
    for f in candidates_tsv:
        t = CandidateTSV(f, num_anns=num_channels - 26)
        counts.append((f, len(t)))
        t.close()
    # the records of the called variants (call.py:308-333) are built by a pool of worker processes WHILE later candidates are
    # scored; for a large call the workers are started here, so that they load while the model is set up
    records = RecordPool(ref_file, chroms, vartype_classes, num_threads, with_pred=False)
    if sum(c[1] for c in counts) >= PREFORK_CANDIDATES:
        records.prefork()
    n_dev = max(1, torch.cuda.device_count())
    if n_dev > 1:
        logger.info("We use {} GPUs!".format(n_dev))
    scorers = []
    try:
        scorers = [CandidateScorer(ck, ensemble=ensemble, device=d) for d in range(n_dev)]
        logger.info("tag: {}".format(scorers[0].meta["tag"]))
        logger.info("model on {} GPU(s) after {:.3f} s".format(n_dev, time.perf_counter() - t_start))
        if not os.path.exists(out_dir):
            os.mkdir(out_dir)
        all_vcf_records, all_vcf_records_none = [], []
        slice_size = int(min(65536, max(batch_size, 16384)))
        handles = []
        with records:
            # the reference's groups (call.py:468-507), in its order; consecutive groups share one run of the pipeline until
            # the run holds RUN_CANDIDATES: a later candidate still replaces an earlier one with the same tag, exactly as the
            # per-group dictionaries merged by dict() do, and no run pays its fill / drain for a few thousand candidates
            run, run_n = [], 0
            groups = plan_call_groups(counts, max_load_candidates)
            for gi, group in enumerate(groups):
                logger.info("Run for candidate files: {}".format(sorted({g[0] for g in group})))
                n_group = sum(g[2] for g in group)
                logger.info("N_dataset: {}".format(n_group))
                if n_group == 0:
                    logger.warning("Skip {} with 0 candidates".format(group[-1][0]))
                else:
                    run += group
                    run_n += n_group
                if run and (run_n >= RUN_CANDIDATES or gi == len(groups) - 1):
                    final_preds, none_preds, true_path = PredTable(vartype_classes), PredTable(vartype_classes), {}
                    t0 = time.perf_counter()
                    run_tsv_pipeline(scorers, run, slice_size, num_threads, final_preds, none_preds, true_path)
                    t1 = time.perf_counter()
                    logger.info("Called {} candidates in this batch.".format(run_n))
                    handles.append(records.submit(final_preds, true_path))
                    t2 = time.perf_counter()
                    all_vcf_records_none.extend(pred_vcf_records_none(none_preds, chroms, with_pred=False))
                    logger.info("{} candidates scored in {:.3f} s; {} called variants handed to the VCF stage in {:.3f} s; "
                                "none.vcf filter {:.3f} s".format(run_n, t1 - t0, len(true_path), t2 - t1, time.perf_counter() - t2))
                    run, run_n = [], 0
            t0 = time.perf_counter()
            for h in handles:
                all_vcf_records.extend(records.collect(h))
            logger.info("VCF records collected after {:.3f} s more".format(time.perf_counter() - t0))
    except BaseException:
        records.close(kill=True)
        raise
    finally:
        for sc in scorers:
            sc.engine.close()
    logger.info("all groups done after {:.3f} s".format(time.perf_counter() - t_start))
    all_vcf_records = dict(all_vcf_records)
    all_vcf_records_none = dict(all_vcf_records_none)
    logger.info("Prepare Output VCF")
    output_vcf = "{}/pred.vcf".format(out_dir)
    write_vcf(get_vcf_records(all_vcf_records), output_vcf, chroms_order, pass_threshold, lowqual_threshold)
    logger.info("Prepare Non-Somatics VCF")
    write_vcf(get_vcf_records(all_vcf_records_none), "{}/none.vcf".format(out_dir), chroms_order, pass_threshold,
              lowqual_threshold)
    logger.info("Calling is Done. ({:.3f} s)".format(time.perf_counter() - t_start))
    return output_vcf
