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
"""
from __future__ import annotations

import gc
import logging

import numpy as np
import torch

from .engine import Engine, VARTYPE_CLASSES, anns_to_255

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


def call_variants(net, vartype_classes, call_loader, out_dir, model_tag, use_cuda):
    """call.py:54-118.  ``net`` is a ``neusomatic_b200.network.NeuSomaticNet`` (possibly wrapped
    in nn.DataParallel).  ``true_path`` maps each non-NONE candidate to its in-memory
    ``non_transformed[:, :, 0:3]`` uint8 matrix (the reference writes it to a PNG, a lossless
    round trip: call.py:89-95,128)."""
    if not use_cuda:
        raise RuntimeError("neusomatic_b200 scores on CUDA only")
    net.eval()
    final_preds, none_preds, true_path = {}, {}, {}
    iii = j = 0
    for data in call_loader:
        (matrices, labels, var_pos_s, var_len_s, non_transformed_matrices), (paths) = data
        iii += 1
        j += len(paths[0])
        matrices = matrices.cuda(non_blocking=True)
        with torch.no_grad():
            outputs, _ = net(matrices)
        o1, o2, o3 = outputs
        # one D2H per batch; softmax / argmax batched on the device
        packed = torch.cat([o1, o2, o3, torch.softmax(o1, 1), torch.softmax(o3, 1),
                            torch.max(o1, 1)[1].float()[:, None], torch.max(o3, 1)[1].float()[:, None]], 1).cpu().numpy()
        n_before = len(final_preds)
        build_pred_dicts(packed[:, 0:4], packed[:, 4:5], packed[:, 5:9], packed[:, 9:13], packed[:, 13:17],
                         packed[:, 17].astype(np.int64), packed[:, 18].astype(np.int64), paths[0],
                         vartype_classes, final_preds, none_preds)
        if len(final_preds) != n_before:
            nt = np.asarray(non_transformed_matrices)
            for i, path_ in enumerate(paths[0]):
                path = path_.split("/")[-1]
                if path in final_preds and path not in true_path:
                    true_path[path] = np.array(nt[i, :, :, 0:3])
        if iii % 10 == 0:
            logger.info("Called {} candidates in this batch.".format(j))
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

    def call_tsv(self, path, batch=65536, num_threads=0, final_preds=None, none_preds=None):
        """call.py:504-534 for one candidates_*.tsv: C++ decode (csrc/nsc_tsv.cpp) -> nsc_score_host_u8 ->
        (final_preds, none_preds), in slices of ``batch`` candidates."""
        from .tsv import CandidateTSV
        tsv = CandidateTSV(path, num_anns=self.C - 26, num_threads=num_threads)
        final_preds = {} if final_preds is None else final_preds
        none_preds = {} if none_preds is None else none_preds
        for first in range(0, len(tsv), batch):
            cand = tsv.decode(first, min(batch, len(tsv) - first))
            self.call(cand, final_preds, none_preds)
        tsv.close()
        return final_preds, none_preds
