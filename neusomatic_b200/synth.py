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


def structured_pileups(n, seed=1234, ensemble=False):
    """S2: one-hot reference row, tumor/normal count rows with a planted variant at ``center``.

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
    anns = None
    if ensemble:
        a = rng.random(size=(n, NUM_ANNS)) * (rng.random(size=(n, NUM_ANNS)) < 0.7)
        anns = np.round(a, 4).astype(np.float32)
    return Candidates(mat, center, tcov, ncov, anns,
                      _tags(rng, vtype, center, length, tcov, ncov))


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
