"""CPU emulation: conv products as  x16*w16 + x8(e4m3)*wlo8(e5m2) + xlo8(e5m2)*w8(e4m3)  (fp8 correction terms)
vs the current  x16*w16 + x16*wlo16 + xlo16*w16, against the golden reference outputs."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from oracle import nsnet_oracle as oracle

NSBLOCKS = [(3, 5, 1, 1, 1), (3, 5, 1, 1, 2), (3, 5, 2, 1, 2), (3, 5, 4, 2, 2)]
f16 = lambda t: t.to(torch.float16).to(torch.float64)
e4 = lambda t: t.to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)
e5 = lambda t: t.to(torch.float32).to(torch.float8_e5m2).to(torch.float64)

def fold(p, conv, bn):
    a = p[bn + ".weight"].double() / torch.sqrt(p[bn + ".running_var"].double() + 1e-5)
    w = p[conv + ".weight"].double() * a[:, None, None, None]
    b = (p[conv + ".bias"].double() - p[bn + ".running_mean"].double()) * a + p[bn + ".bias"].double()
    return w, b

def conv(x, w, b, mode, **kw):
    if mode == "fp32":
        return F.conv2d(x, w, b, **kw)
    xh = f16(x); xl = x - xh
    wh = f16(w); wl = w - wh
    if mode == "split16":
        return F.conv2d(xh, wh, b, **kw) + F.conv2d(xh, f16(wl), None, **kw) + F.conv2d(f16(xl), wh, None, **kw)
    if mode == "fp8corr":
        return F.conv2d(xh, wh, b, **kw) + F.conv2d(e4(x), e5(wl), None, **kw) + F.conv2d(e5(xl), e4(w), None, **kw)
    if mode == "fp8corr_e4lo":   # lo parts scaled into e4m3 range would need separate accumulators; reference point
        return F.conv2d(xh, wh, b, **kw) + F.conv2d(e4(x), wl, None, **kw) + F.conv2d(xl, e4(w), None, **kw)
    if mode == "single16":
        return F.conv2d(xh, wh, b, **kw)

def forward(p, x, mode):
    x = x.double()
    w, b = fold(p, "conv1", "bn1")
    x = F.max_pool2d(F.relu(conv(x, w, b, mode, padding=(0, 1))), (1, 3), stride=(1, 1), padding=(0, 1))
    for i, (k1, k2, d1, d2, st) in enumerate(NSBLOCKS):
        pre = "res_layers.%d." % i
        w, b = fold(p, pre + "conv_r1", pre + "bn_r1")
        y1 = F.relu(conv(x, w, b, mode, dilation=d1, padding=(d1 * (k1 - 1)) // 2))
        w, b = fold(p, pre + "conv_r2", pre + "bn_r2")
        y2 = conv(y1, w, b, mode, dilation=d2, padding=(d2 * (k2 - 1)) // 2)
        x = F.max_pool2d(x + y2, (1, 3), stride=(1, st), padding=(0, 1))
    x3 = F.relu(F.linear(x.reshape(x.shape[0], -1), p["fc1.weight"].double(), p["fc1.bias"].double()))
    return [F.linear(x3, p[n + ".weight"].double(), p[n + ".bias"].double()).float().numpy() for n in ("fc2", "fc3", "fc4")]

torch.set_num_threads(8)
for name in ("std_v010", "ens_v010", "std_v014"):
    g, ck = load_golden(name)
    p = {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}
    x = torch.from_numpy(oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], g.get("anns"), int(ck["coverage_thr"])))
    for mode in ("fp32", "split16", "fp8corr", "single16"):
        o1, o2, o3 = forward(p, x, mode)
        r1, r3 = oracle.softmax(g["o1"]), oracle.softmax(g["o3"])
        print("%s %-9s max|dlogit|=%.2e max|dprob|=%.2e max|dpos|=%.2e flips=%d/%d" % (
            name, mode, max(np.abs(o1 - g["o1"]).max(), np.abs(o3 - g["o3"]).max()),
            max(np.abs(oracle.softmax(o1) - r1).max(), np.abs(oracle.softmax(o3) - r3).max()), np.abs(o2 - g["o2"]).max(),
            (o1.argmax(1) != g["o1"].argmax(1)).sum(), (o3.argmax(1) != g["o3"].argmax(1)).sum()))
