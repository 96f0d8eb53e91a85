"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

numpy restatement of the reference's candidate-scoring hot path:

  * ``assemble_channels``   -- NeuSomaticDataset.__getitem__, is_test branch
                               (reference neusomatic/python/dataloader.py:383-417 plus the
                               matrix_transform at dataloader.py:25-36 / call.py:408)
  * ``forward``             -- NeuSomaticNet.forward (network.py:66-78, NSBlock 31-36)
  * ``call_variants``       -- the per-batch result extraction of call.py:80-114
  * ``decode_tsv_line`` / ``encode_tsv_line`` -- candidate TSV format
                               (dataloader.py:38-58, generate_dataset.py:654-657,1569-1581)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm
may import this module.  The arithmetic of the reference lives in PyTorch CPU operators
(conv2d / batch_norm / max_pool2d / linear, README.md:43 pins torch 1.1.0); this file restates
them as zero-padded shifted-window matmuls in numpy float32 (float64 optional for a noise
floor).  Parity pin: ``tests/golden/*.npz`` hold outputs of the unmodified reference
``network.py`` + ``dataloader.py`` (imported from /root/reference in the build container by
``tests/golden/make_golden.py``) on seeded inputs; ``tests/test_oracle.py`` checks this file
against them.
"""
from __future__ import annotations

import base64
import zlib

import numpy as np

VARTYPE_CLASSES = ["DEL", "INS", "NONE", "SNP"]          # call.py:407
# (ks_1, ks_2, dl_1, dl_2, mp_ks, mp_st) per NSBlock       network.py:48-53
NSBLOCKS = [(3, 5, 1, 1, 3, 1), (3, 5, 1, 1, 3, 2), (3, 5, 2, 1, 3, 2), (3, 5, 4, 2, 3, 2)]
BN_EPS = 1e-5                                             # nn.BatchNorm2d default


# ----------------------------------------------------------------------------------------
# state_dict handling (call.py:441-460: "module." prefix reconciliation)
# ----------------------------------------------------------------------------------------
def normalize_state_dict(sd, dtype=np.float32):
    """Strip a ``module.`` prefix, drop ``num_batches_tracked`` and convert to numpy."""
    out = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if k.endswith("num_batches_tracked"):
            continue
        out[k] = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v).astype(dtype)
    return out


# ----------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------
def conv2d(x, w, b, dil, pad):
    """x [B,Ci,H,W], w [Co,Ci,kh,kw]; stride 1, zero padding ``pad`` = (ph, pw), dilation dil."""
    B, Ci, Hh, Ww = x.shape
    Co, _, kh, kw = w.shape
    ph, pw = pad
    xp = np.zeros((B, Ci, Hh + 2 * ph, Ww + 2 * pw), x.dtype)
    xp[:, :, ph:ph + Hh, pw:pw + Ww] = x
    out = np.zeros((B, Co, Hh * Ww), x.dtype)
    for dy in range(kh):
        for dx in range(kw):
            xs = xp[:, :, dy * dil:dy * dil + Hh, dx * dil:dx * dil + Ww].reshape(B, Ci, Hh * Ww)
            out += np.matmul(w[:, :, dy, dx][None], xs)
    out += b[None, :, None]
    return out.reshape(B, Co, Hh, Ww)


def batchnorm_eval(x, p, prefix):
    """Inference BatchNorm2d: y = (x - mean) / sqrt(var + eps) * gamma + beta."""
    dt = x.dtype.type
    invstd = dt(1.0) / np.sqrt(p[prefix + ".running_var"] + dt(BN_EPS))
    alpha = p[prefix + ".weight"] * invstd
    beta = p[prefix + ".bias"] - p[prefix + ".running_mean"] * alpha
    return x * alpha[None, :, None, None] + beta[None, :, None, None]


def maxpool_1x3(x, stride):
    """MaxPool2d((1,3), padding=(0,1), stride=(1,stride)); padding value is -inf."""
    B, Cc, Hh, Ww = x.shape
    xp = np.full((B, Cc, Hh, Ww + 2), -np.inf, x.dtype)
    xp[..., 1:Ww + 1] = x
    wo = (Ww + 2 - 3) // stride + 1
    idx = np.arange(wo) * stride
    return np.maximum(np.maximum(xp[..., idx], xp[..., idx + 1]), xp[..., idx + 2])


def nsblock(x, p, prefix, ks1, ks2, dl1, dl2, mp_st):
    """network.py:31-36.  Note: no ReLU after the residual add."""
    y1 = conv2d(x, p[prefix + "conv_r1.weight"], p[prefix + "conv_r1.bias"], dl1,
                ((dl1 * (ks1 - 1)) // 2,) * 2)
    y1 = np.maximum(batchnorm_eval(y1, p, prefix + "bn_r1"), 0)
    y2 = conv2d(y1, p[prefix + "conv_r2.weight"], p[prefix + "conv_r2.bias"], dl2,
                ((dl2 * (ks2 - 1)) // 2,) * 2)
    y2 = batchnorm_eval(y2, p, prefix + "bn_r2")
    return maxpool_1x3(x + y2, mp_st)


def forward(p, x, return_internals=False):
    """NeuSomaticNet.forward (network.py:66-78) on x [B,C,5,32]; p from normalize_state_dict.

    Returns (type_logits [B,4], pos [B,1], len_logits [B,4]) (+ internal_outs if requested).
    """
    x = np.ascontiguousarray(x, dtype=p["conv1.weight"].dtype)
    y = conv2d(x, p["conv1.weight"], p["conv1.bias"], 1, (0, 1))
    y = maxpool_1x3(np.maximum(batchnorm_eval(y, p, "bn1"), 0), 1)
    internals = [y]
    for i, (ks1, ks2, dl1, dl2, _mp_ks, mp_st) in enumerate(NSBLOCKS):
        y = nsblock(y, p, "res_layers.%d." % i, ks1, ks2, dl1, dl2, mp_st)
    internals.append(y)
    x2 = y.reshape(y.shape[0], -1)                       # index c*20 + h*4 + w
    x3 = np.maximum(x2 @ p["fc1.weight"].T + p["fc1.bias"], 0)
    internals.extend([x2, x3])
    o1 = x3 @ p["fc2.weight"].T + p["fc2.bias"]
    o2 = x3 @ p["fc3.weight"].T + p["fc3.bias"]
    o3 = x3 @ p["fc4.weight"].T + p["fc4.bias"]
    if return_internals:
        return (o1, o2, o3), internals
    return o1, o2, o3


def softmax(z):
    """F.softmax over the last axis, float32 in -> float32 out."""
    z = z - z.max(-1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(-1, keepdims=True)).astype(z.dtype)


# ----------------------------------------------------------------------------------------
# dataloader restatement (is_test branch)
# ----------------------------------------------------------------------------------------
def parse_tag(tag):
    """dataloader.py:267-274: tag = chrom.pos.ref.alt.type.center.len.tumor_cov.normal_cov."""
    _, _, _, _, vartype, center, length, tumor_cov, normal_cov = tag.split("/")[-1].split(".")
    return vartype, int(center), int(length), int(tumor_cov), int(normal_cov)


def assemble_channels(matrices, center, tumor_cov, normal_cov, anns, coverage_thr,
                      normalize_channels=False):
    """uint8 [N,5,32,23] (+ tag fields, annotations) -> float32 [N,C,5,32], C = 26 + n_anns.

    Follows dataloader.py:383-411 in float64, then ``.float().div(255)`` in float32 and the
    matrix_transform((.5,.5,.5),(.5,.5,.5)) that touches **channels 0..2 only**
    (zip over 3-tuples, dataloader.py:33-35).
    """
    n = matrices.shape[0]
    na = 0 if anns is None else anns.shape[1]
    m = np.zeros((n, 5, 32, 26 + na), np.float64)
    m[..., 0:23] = matrices
    if normalize_channels:                                        # dataloader.py:386-388
        m[..., 3:23:2] *= (m[..., 1:2] / 255.0)
        m[..., 4:23:2] *= (m[..., 2:3] / 255.0)
    mx = m[..., 0].reshape(n, -1).max(1)                          # np.max(matrix[:, :, 0])
    for i in range(n):
        m[i, :, int(center[i]), 23] = mx[i]                       # dataloader.py:390
    thr = float(coverage_thr)
    m[..., 24] = (np.minimum(tumor_cov, coverage_thr) / thr * 255.0)[:, None, None]
    m[..., 25] = (np.minimum(normal_cov, coverage_thr) / thr * 255.0)[:, None, None]
    if na:
        m[..., 26:] = (np.asarray(anns, np.float64) * 255.0)[:, None, None, :]
    x = m.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)
    x[:, 0:3] = (x[:, 0:3] - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x)


def non_transformed(matrices, center):
    """dataloader.py:398-404: uint8 [N,5,32,3] = (ch0, ch1, center marker)."""
    n = matrices.shape[0]
    out = np.zeros((n, 5, 32, 3), np.uint8)
    out[..., 0:2] = matrices[..., 0:2]
    mx = matrices[..., 0].reshape(n, -1).max(1)
    for i in range(n):
        out[i, :, int(center[i]), 2] = mx[i]
    return out


def encode_tsv_line(i, tag, matrix, anns=None):
    """generate_dataset.py:654-657, 1569-1581."""
    b64 = base64.b64encode(zlib.compress(np.ascontiguousarray(matrix, np.uint8))).decode("ascii")
    fields = [str(i), "1", tag, b64]
    if anns is not None and len(anns):
        fields += [str(float(a)) for a in anns]
    return "\t".join(fields) + "\n"


def decode_tsv_line(line):
    """dataloader.py:38-58 (np.fromstring replaced by its numpy>=2 spelling np.frombuffer)."""
    fields = line.strip().split()
    tag = fields[2]
    im = np.frombuffer(zlib.decompress(base64.b64decode(fields[3])), np.uint8).reshape(5, 32, 23)
    anns = [float(a) for a in fields[4:]]
    return tag, im, anns


# ----------------------------------------------------------------------------------------
# call_variants restatement (call.py:80-114)
# ----------------------------------------------------------------------------------------
def call_variants(o1, o2, o3, tags):
    """Returns (final_preds, none_preds) exactly as call.py builds them."""
    o1 = np.asarray(o1, np.float32)
    o2 = np.asarray(o2, np.float32)
    o3 = np.asarray(o3, np.float32)
    predicted = o1.argmax(1)                                      # torch.max: first max
    len_pred = o3.argmax(1)
    p1 = softmax(o1)
    p3 = softmax(o3)
    final_preds, none_preds = {}, {}
    for i, path_ in enumerate(tags):
        path = path_.split("/")[-1]
        entry = [VARTYPE_CLASSES[predicted[i]], o2[i], len_pred[i],
                 [round(v, 4) for v in p1[i]], [round(v, 4) for v in p3[i]],
                 [round(v, 4) for v in o1[i]], [round(v, 4) for v in o3[i]]]
        if VARTYPE_CLASSES[predicted[i]] != "NONE":
            final_preds[path] = entry
        else:
            none_preds[path] = entry
    return final_preds, none_preds
