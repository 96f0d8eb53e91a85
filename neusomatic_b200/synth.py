"""Seeded synthetic candidate generators (numpy only, platform independent).

These produce inputs *of the shape and semantics* the reference stores on disk
(``uint8[5,32,23]`` pileup images + tag fields + optional 93 ensemble annotations;
reference: neusomatic/python/generate_dataset.py:598-657 for the channel meaning and
the tag format, dataloader.py:245-282 for how the tag is consumed).  They are used by
the golden-vector script, the parity tests and ``bench.py``.

Three families:
  S1 ``uniform_pileups``    -- masked uniform-random bytes (throughput only).
  S2 ``structured_pileups`` -- planted SNP/INS/DEL pileups (parity-sensitive).
  S3 is S2 pushed through a randomly initialised net (see ``random_state_dict``).
"""
from __future__ import annotations

import numpy as np

VARTYPE_CLASSES = ["DEL", "INS", "NONE", "SNP"]  # call.py:407
NUM_ANNS = 93
H, W, CH = 5, 32, 23


class Candidates:
    """A batch of compact candidates (what one TSV shard decodes to)."""

    def __init__(self, matrices, center, tumor_cov, normal_cov, anns, tags):
        self.matrices = matrices      # uint8 [N,5,32,23]
        self.center = center          # int32 [N]
        self.tumor_cov = tumor_cov    # int32 [N]
        self.normal_cov = normal_cov  # int32 [N]
        self.anns = anns              # float32 [N,93] or None
        self.tags = tags              # list[str]

    def __len__(self):
        return self.matrices.shape[0]

    @property
    def meta(self):
        """int32 [N,3] = (center, tumor_cov, normal_cov): the C-ABI's ``meta`` argument."""
        return np.stack([self.center, self.tumor_cov, self.normal_cov], 1).astype(np.int32)


def _tags(rng, vtype, center, length, tcov, ncov):
    n = len(vtype)
    pos = rng.integers(1, 50_000_000, size=n)
    bases = np.array(list("ACGT"))
    ref = bases[rng.integers(0, 4, size=n)]
    alt = bases[rng.integers(0, 4, size=n)]
    return ["{}.{}.{}.{}.{}.{}.{}.{}.{}".format(
        int(21), int(pos[i]), ref[i], alt[i], VARTYPE_CLASSES[int(vtype[i])],
        int(center[i]), int(length[i]), int(tcov[i]), int(ncov[i])) for i in range(n)]


def uniform_pileups(n, seed=1234, ensemble=False):
    """S1: Bernoulli(0.4)-masked uniform bytes; saturates the net, fine for throughput."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = rng.integers(0, 256, size=(n, H, W, CH), dtype=np.uint8)
    m *= (rng.random(size=(n, H, W, CH)) < 0.4).astype(np.uint8)
    center = rng.integers(2, 30, size=n).astype(np.int32)
    tcov = rng.integers(5, 401, size=n).astype(np.int32)
    ncov = rng.integers(5, 401, size=n).astype(np.int32)
    anns = np.round(rng.random(size=(n, NUM_ANNS)), 4).astype(np.float32) if ensemble else None
    vtype = rng.integers(0, 4, size=n)
    length = rng.integers(0, 4, size=n)
    return Candidates(m, center, tcov, ncov, anns, _tags(rng, vtype, center, length, tcov, ncov))


def _structured(n, seed, ensemble, window=None):
    """S2 (see structured_pileups); also returns the planted layout for genome_candidates.

    ``window``: None = all 32 columns populated (the golden parity sets); an int = the scanner's layout
    (generate_dataset.py:452-458: 2 * matrix_base_pad + 1 = 15 reference columns plus the insertion columns, centred in the
    32-column matrix, empty columns either side, the variant in the middle of the window).

    S2: one-hot reference row, tumor/normal count rows with a planted variant at ``center``.

    Channel semantics follow generate_dataset.py:598-653: 0 reference, 1/2 tumor/normal base
    counts, 3/4 mean BQ, 5/6 mean MQ, 7/8 forward strand, 9-12 clips, 13-22 five tag means;
    auxiliary channels are non-zero only where the matching count is non-zero.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    ar = np.arange(n)
    refbase = rng.integers(1, 5, size=(n, W))                 # rows 1..4 = A,C,G,T; row 0 = gap
    center = rng.integers(7, 25, size=n).astype(np.int32)
    vtype = rng.choice(4, size=n, p=[0.15, 0.15, 0.35, 0.35])  # DEL, INS, NONE, SNP
    length = np.where(vtype == 2, 0, np.where(vtype == 3, 1, rng.integers(1, 5, size=n)))
    colmask = None
    if window is not None:
        ncols = window + np.where(vtype == 1, length, 0)
        off = (W - ncols) // 2
        center = (off + window // 2).astype(np.int32)
        colmask = (np.arange(W)[None, :] >= off[:, None]) & (np.arange(W)[None, :] < (off + ncols)[:, None])
    tcov = rng.integers(10, 401, size=n).astype(np.int32)
    ncov = rng.integers(10, 401, size=n).astype(np.int32)
    af_t = np.where(vtype == 2, rng.uniform(0.0, 0.02, size=n), rng.uniform(0.03, 0.6, size=n))
    af_n = rng.uniform(0.0, 0.015, size=n)

    cols = np.arange(W)[None, :]
    span = (cols >= center[:, None]) & (cols < (center + np.maximum(length, 1))[:, None])  # [n,W]
    is_ins = (vtype == 1)[:, None] & span
    is_del = (vtype == 0)[:, None] & span
    is_snp = (vtype == 3)[:, None] & (cols == center[:, None])
    # insertion columns have a gap in the reference
    refrow = np.where(is_ins, 0, refbase)
    altrow = np.where(is_del, 0, (refbase + rng.integers(0, 3, size=(n, W))) % 4 + 1)
    altrow = np.where(altrow == refrow, altrow % 4 + 1, altrow)
    var_col = is_ins | is_del | is_snp

    def counts(cov, af):
        depth = cov[:, None] * rng.uniform(0.7, 1.0, size=(n, W))
        alt_n = np.where(var_col, depth * af[:, None], 0.0)
        c = np.zeros((n, H, W), np.float64)
        noise = rng.random(size=(n, H, W)) < 0.04
        c += noise * depth[:, None, :] * rng.uniform(0.0, 0.02, size=(n, H, W))
        np.put_along_axis(c, refrow[:, None, :], (depth - alt_n)[:, None, :], 1)
        cur = np.take_along_axis(c, altrow[:, None, :], 1)
        np.put_along_axis(c, altrow[:, None, :], cur + alt_n[:, None, :], 1)
        return np.round(c)

    ct = counts(tcov, af_t)
    cn = counts(ncov, af_n)
    m = np.zeros((n, H, W, CH), np.float64)
    ref1h = np.zeros((n, H, W), np.float64)
    np.put_along_axis(ref1h, refrow[:, None, :], 1.0, 1)
    tmax = ct.reshape(n, -1).max(1)[:, None, None] + 1e-5
    nmax = cn.reshape(n, -1).max(1)[:, None, None] + 1e-5
    m[..., 0] = ref1h * (np.maximum(tcov[:, None, None], 1) /
                         np.maximum(tcov[:, None, None], tmax)) * 255
    m[..., 1] = ct / tmax * 255
    m[..., 2] = cn / nmax * 255
    for k, (c, mx) in enumerate(((ct, tmax), (cn, nmax))):
        on = c > 0
        m[..., 3 + k] = on * rng.uniform(22, 41, size=(n, H, W)) / 41.0 * 255
        m[..., 5 + k] = on * rng.uniform(35, 60, size=(n, H, W)) / 70.0 * 255
        m[..., 7 + k] = np.round(c * rng.uniform(0.3, 0.7, size=(n, H, W))) / mx * 255
        m[..., 9 + k] = np.round(c * (rng.random(size=(n, H, W)) < 0.1) * 0.1) / mx * 255
        m[..., 11 + k] = np.round(c * (rng.random(size=(n, H, W)) < 0.1) * 0.1) / mx * 255
        for t, (lo, hi) in enumerate(((0, 6), (80, 100), (0, 60), (0, 100), (0, 12))):
            m[..., 13 + 2 * t + k] = on * rng.uniform(lo, hi, size=(n, H, W)) / 100.0 * 255
    mat = np.clip(m, 0, 255).astype(np.uint8)
    if colmask is not None:
        mat *= colmask[:, None, :, None].astype(np.uint8)
        refrow = np.where(colmask, refrow, -1)              # -1: empty column (neither gap nor base)
    anns = None
    if ensemble:
        a = rng.random(size=(n, NUM_ANNS)) * (rng.random(size=(n, NUM_ANNS)) < 0.7)
        anns = np.round(a, 4).astype(np.float32)
    layout = dict(vtype=vtype, length=length, refrow=refrow, altrow=altrow)
    return rng, mat, center, tcov, ncov, anns, layout


def structured_pileups(n, seed=1234, ensemble=False):
    """S2: planted SNP/INS/DEL pileups with random (inconsistent) chrom/pos/ref/alt tag fields."""
    rng, mat, center, tcov, ncov, anns, lay = _structured(n, seed, ensemble)
    return Candidates(mat, center, tcov, ncov, anns,
                      _tags(rng, lay["vtype"], center, lay["length"], tcov, ncov))


GENOME_CHROMS = ["chr1", "chr2", "chrX"]


def genome_candidates(n, seed=1234, ensemble=False, window=15):
    """S2 pileups whose tags AND a synthetic reference genome are consistent with the matrices, so that the
    reference's VCF stage (call.py:121-305) accepts them: the non-gap reference columns of candidate i map to
    consecutive positions of its own 1 kb region of chromosome ``i % 3`` and the tag carries the pos/ref/alt the
    scanner would have written (generate_dataset.py:598-657 anchors: SNP column = pos, DEL column = pos + 1,
    INS column - 1 = pos).  Returns (Candidates, {chrom name: sequence}, chrom names)."""
    rng, mat, center, tcov, ncov, anns, lay = _structured(n, seed, ensemble, window)
    g = np.random.Generator(np.random.PCG64(seed + 7919))
    acgt = "ACGT"
    per = (n + len(GENOME_CHROMS) - 1) // len(GENOME_CHROMS)
    seqs = [list(np.array(list(acgt))[g.integers(0, 4, size=1000 * per + 200)]) for _ in GENOME_CHROMS]
    tags = []
    for i in range(n):
        ci, g0 = i % len(GENOME_CHROMS), 1000 * (i // len(GENOME_CHROMS)) + 101      # 1-based position of the first column
        refrow, altrow = lay["refrow"][i], lay["altrow"][i]
        nz = [c for c in range(W) if refrow[c] > 0]
        col2pos = {c: g0 + j for j, c in enumerate(nz)}
        for c in nz:
            seqs[ci][col2pos[c] - 1] = acgt[refrow[c] - 1]
        vt, ce, ln = int(lay["vtype"][i]), int(center[i]), int(lay["length"][i])
        base = lambda pos: seqs[ci][pos - 1]
        if vt == 0:                                   # DEL: ref = anchor base + deleted bases
            pos = col2pos[ce] - 1
            ref = "".join(base(pos + k) for k in range(ln + 1))
            alt = base(pos)
        elif vt == 1:                                 # INS: alt = anchor base + inserted bases
            pos = col2pos[ce - 1]
            ref = base(pos)
            alt = ref + "".join(acgt[max(int(altrow[ce + k]), 1) - 1] for k in range(ln))
        else:                                         # SNP / NONE
            pos = col2pos[ce]
            ref = base(pos)
            alt = acgt[int(altrow[ce]) - 1] if altrow[ce] > 0 else acgt[(acgt.index(ref) + 1) % 4]
        tags.append("{}.{}.{}.{}.{}.{}.{}.{}.{}".format(ci, pos, ref, alt, VARTYPE_CLASSES[vt], ce, ln,
                                                        int(tcov[i]), int(ncov[i])))
    fasta = {name: "".join(sq) for name, sq in zip(GENOME_CHROMS, seqs)}
    return Candidates(mat, center, tcov, ncov, anns, tags), fasta, list(GENOME_CHROMS)




def random_state_dict(num_channels, seed=0):
    """S3 weights: seeded randomly initialised NeuSomaticNet parameters (numpy, fp32) with
    randomised BN running statistics, in the reference's state_dict naming (network.py:41-64)."""
    rng = np.random.Generator(np.random.PCG64(seed))

    def conv(co, ci, kh, kw):
        bound = 1.0 / np.sqrt(ci * kh * kw)
        return (rng.uniform(-bound, bound, size=(co, ci, kh, kw)).astype(np.float32),
                rng.uniform(-bound, bound, size=(co,)).astype(np.float32))

    def bn(prefix, sd):
        sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, size=64).astype(np.float32)
        sd[prefix + ".bias"] = rng.uniform(-0.3, 0.3, size=64).astype(np.float32)
        sd[prefix + ".running_mean"] = rng.uniform(-0.3, 0.3, size=64).astype(np.float32)
        sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, size=64).astype(np.float32)

    sd = {}
    sd["conv1.weight"], sd["conv1.bias"] = conv(64, num_channels, 1, 3)
    bn("bn1", sd)
    for i, (k1, k2) in enumerate(((3, 5),) * 4):
        p = "res_layers.%d." % i
        sd[p + "conv_r1.weight"], sd[p + "conv_r1.bias"] = conv(64, 64, k1, k1)
        bn(p + "bn_r1", sd)
        sd[p + "conv_r2.weight"], sd[p + "conv_r2.bias"] = conv(64, 64, k2, k2)
        bn(p + "bn_r2", sd)
    for name, (o, i) in (("fc1", (240, 1280)), ("fc2", (4, 240)), ("fc3", (1, 240)), ("fc4", (4, 240))):
        bound = 1.0 / np.sqrt(i)
        sd[name + ".weight"] = rng.uniform(-bound, bound, size=(o, i)).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-bound, bound, size=(o,)).astype(np.float32)
    return sd
