"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU emulation (float64 accumulation) of the operand arithmetic of the tensor-core modes of
libneusomatic_cuda (csrc/nsc_umma.cu), used to choose and pin the precision design:

  split16 : x*w ~ x16*w16 + x16*fp16(w_lo) + fp16(x_lo)*w16                        (NSC_PREC_SPLIT16)
  split8  : x*w ~ x16*w16 + [e4m3(x)*e4m3(w_lo*2^15) + e4m3(x_lo*2^12)*e4m3(w*2^3)] * 2^-15   (NSC_PREC_SPLIT8)
  single16: x16*w16                                                                 (NSC_PREC_SINGLE16)

with x16 = fp16(x), x_lo = x - x16 (same for w), BatchNorm folded into the weights in float64 as
nsc_api.cu:fold_conv_bn does, conv1 and the linear layers as the library runs them (conv1 split16 in
every tensor mode).  Network structure: /root/reference/neusomatic/python/network.py:17-36, 41-78.
"""
import torch
import torch.nn.functional as F

NSBLOCKS = [(3, 5, 1, 1, 1), (3, 5, 1, 1, 2), (3, 5, 2, 1, 2), (3, 5, 4, 2, 2)]   # network.py:48-53


def _f16(t):
    return t.to(torch.float16).to(torch.float64)


def _e4(t):
    return t.clamp(-448.0, 448.0).to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)


def fold(p, conv, bn, eps=1e-5):
    a = p[bn + ".weight"].double() / torch.sqrt(p[bn + ".running_var"].double() + eps)
    w = p[conv + ".weight"].double() * a[:, None, None, None]
    b = (p[conv + ".bias"].double() - p[bn + ".running_mean"].double()) * a + p[bn + ".bias"].double()
    return w.float().double(), b.float().double()      # the library keeps the folded parameters in fp32


def conv(x, w, b, mode, **kw):
    if mode == "exact":
        return F.conv2d(x, w, b, **kw)
    xh, wh = _f16(x), _f16(w)
    xl, wl = x - xh, w - wh
    if mode == "single16":
        return F.conv2d(xh, wh, b, **kw)
    if mode == "split16":
        return F.conv2d(xh, wh, b, **kw) + F.conv2d(xh, _f16(wl), None, **kw) + F.conv2d(_f16(xl), wh, None, **kw)
    if mode == "split8":
        corr = F.conv2d(_e4(x), _e4(wl * 2.0 ** 15), None, **kw) + F.conv2d(_e4(xl * 2.0 ** 12), _e4(w * 8.0), None, **kw)
        return F.conv2d(xh, wh, b, **kw) + corr / 2.0 ** 15
    raise ValueError(mode)


def forward(params, x, mode):
    """params: state_dict without the `module.` prefix; x: float32 [B,C,5,32].  Returns (o1, o2, o3) as float32 numpy."""
    p = params
    x = x.double()
    first = "split16" if mode == "split8" else mode
    w, b = fold(p, "conv1", "bn1")
    x = F.max_pool2d(F.relu(conv(x, w, b, first, padding=(0, 1))), (1, 3), stride=(1, 1), padding=(0, 1))
    x = x.float().double()
    for i, (k1, k2, d1, d2, st) in enumerate(NSBLOCKS):
        pre = "res_layers.%d." % i
        w, b = fold(p, pre + "conv_r1", pre + "bn_r1")
        y1 = F.relu(conv(x, w, b, mode, dilation=d1, padding=(d1 * (k1 - 1)) // 2)).float().double()
        w, b = fold(p, pre + "conv_r2", pre + "bn_r2")
        y2 = conv(y1, w, b, mode, dilation=d2, padding=(d2 * (k2 - 1)) // 2)
        x = F.max_pool2d(x + y2, (1, 3), stride=(1, st), padding=(0, 1)).float().double()
    x3 = F.relu(F.linear(x.reshape(x.shape[0], -1), p["fc1.weight"].double(), p["fc1.bias"].double()))
    return [F.linear(x3, p[n + ".weight"].double(), p[n + ".bias"].double()).float().numpy() for n in ("fc2", "fc3", "fc4")]
