"""Where does the tensor path's residual error come from?  Run on the GPU box.

    python tools/accum_bias_probe.py [n]

Split-fp16 operands reproduce the fp32 products exactly; what is left (max |d logit| ~5e-4 against 4e-5 for the CUDA-core
fp32 path) is the fp32 accumulation inside the tensor core.  If that accumulation rounds toward zero, the error is not
noise but a systematic shrink of every accumulator, proportional to its value and to the length of the chain -- and a
systematic error can be compensated in the folded weights.  This tool measures it:

  1. per layer output (stem, the four NSBlocks, fc1): the GPU activations (debug copies) against a float64 forward of the
     same candidates -- the least-squares slope of (gpu - ref) on ref over the positive activations (= relative shrink),
     the RMS of what the slope does not explain, and the same for the CUDA-core fp32 mode as the noise floor;
  2. the end-to-end effect of scaling every convolution / fc1 weight by (1 + delta): max |d logit|, max |d softmax| and the
     share of 4-decimal probabilities that differ from the float64 reference, per delta.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neusomatic_b200 import Engine, synth                    # noqa: E402
from neusomatic_b200.call import load_checkpoint             # noqa: E402
from oracle import nsnet_oracle as oracle                    # noqa: E402


def forward64(p, x):
    """float64 forward with every debug point of the library: [stem, block0..3, fc1], logits."""
    y = oracle.conv2d(x, p["conv1.weight"], p["conv1.bias"], 1, (0, 1))
    y = oracle.maxpool_1x3(np.maximum(oracle.batchnorm_eval(y, p, "bn1"), 0), 1)
    pts = [y]
    for i, (k1, k2, d1, d2, _k, st) in enumerate(oracle.NSBLOCKS):
        y = oracle.nsblock(y, p, "res_layers.%d." % i, k1, k2, d1, d2, st)
        pts.append(y)
    x2 = y.reshape(len(y), -1)
    x3 = np.maximum(x2 @ p["fc1.weight"].T + p["fc1.bias"], 0)
    pts.append(x3)
    o1 = x3 @ p["fc2.weight"].T + p["fc2.bias"]
    o2 = x3 @ p["fc3.weight"].T + p["fc3.bias"]
    o3 = x3 @ p["fc4.weight"].T + p["fc4.bias"]
    return pts, (o1, o2, o3)


def fit(g, r):
    m = r > 0.05
    g, r = g[m].astype(np.float64), r[m]
    e = g - r
    slope = float((e * r).sum() / (r * r).sum())
    resid = e - slope * r
    return {"n": int(m.sum()), "slope": slope, "rms_err": float(np.sqrt((e * e).mean())), "rms_resid": float(np.sqrt((resid * resid).mean())),
            "mean_err_over_mean_val": float(e.mean() / r.mean()), "max_abs_err": float(np.abs(e).max())}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    ck = torch.load(os.path.join(ROOT, "tests", "golden", "std_v014.pth"), map_location="cpu", weights_only=True)
    sd, meta = load_checkpoint(ck)
    cand = synth.structured_pileups(n, seed=4242)
    x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, None, int(meta["coverage_thr"]))
    p64 = oracle.normalize_state_dict(sd, np.float64)
    pts, (r1, r2, r3) = forward64(p64, x.astype(np.float64))
    xd = torch.from_numpy(x).cuda()
    names = ["stem", "block0", "block1", "block2", "block3", "fc1"]
    out = {"n": n, "layers": {}, "scaled_weights": []}
    for mode, mname in ((2, "FP32"), (0, "SPLIT16"), (3, "SPLIT8")):
        eng = Engine(sd, 26, precision=mode)
        eng.set_debug(n)
        eng.forward_f32(xd)
        out["layers"][mname] = {nm: fit(eng.internal(i, n).cpu().numpy(), pts[i]) for i, nm in enumerate(names)}
        eng.close()
        print(mname, json.dumps(out["layers"][mname]), flush=True)

    q1, q3 = oracle.softmax(r1), oracle.softmax(r3)

    def score(sdict, mode):
        eng = Engine(sdict, 26, precision=mode)
        o1, o2, o3, p1, p3, a1, a3 = [t.cpu().numpy() for t in eng.forward_f32(xd, extras=True)]
        eng.close()
        mm = np.concatenate([np.round(p1, 4) != np.round(q1.astype(np.float32), 4), np.round(p3, 4) != np.round(q3.astype(np.float32), 4)], 1)
        return {"max_dlogit": float(max(np.abs(o1 - r1).max(), np.abs(o3 - r3).max())), "max_dprob": float(max(np.abs(p1 - q1).max(), np.abs(p3 - q3).max())),
                "max_dpos": float(np.abs(o2 - r2).max()), "mean_signed_dlogit_top": float((o1.max(1) - r1.max(1)).mean()),
                "mismatch_4dp_entries": float(mm.mean())}
    for delta in (0.0, 1e-6, 2e-6, 3e-6, 4e-6, 6e-6, 8e-6, 1.2e-5):
        s2 = dict(sd)
        for k in sd:
            if k.endswith("weight") and ("conv" in k or k == "fc1.weight"):
                s2[k] = (sd[k].double() * (1.0 + delta)).float()
        row = {"delta": delta, "SPLIT16": score(s2, 0), "SPLIT8": score(s2, 3)}
        out["scaled_weights"].append(row)
        print(json.dumps(row), flush=True)
    out["FP32_reference_mode"] = score(sd, 2)
    print(json.dumps(out["FP32_reference_mode"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "accum_bias_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
