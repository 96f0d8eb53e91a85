"""Which layers can take the fp8 correction products?  Run on the GPU box.

    python tools/s8_mask_sweep.py [kappa] [mask ...]

NSC_PREC_SPLIT8 with NSC_S8_MASK = mask (bit l: 64 -> 64 convolution l on the fp8-correction path, else split fp16), the
round-toward-zero compensation at `kappa`: parity metrics on the golden sets / S2 sets (tools/rz_kappa_sweep.py) and the
time of a 32 768-candidate forward pass (CUDA events, inputs resident)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden                               # noqa: E402
from neusomatic_b200 import Engine, synth                     # noqa: E402
from neusomatic_b200.call import load_checkpoint              # noqa: E402
from oracle import nsnet_oracle as oracle, torch_port         # noqa: E402
from rz_kappa_sweep import metrics                             # noqa: E402


def main():
    kappa = float(sys.argv[1]) if len(sys.argv) > 1 else 1.4
    masks = [int(a, 0) for a in sys.argv[2:]] or [0xFF, 0x3F, 0x0F, 0x03, 0x55, 0xAA, 0xF0, 0x00]
    os.environ["NSC_RZ_KAPPA"] = repr(kappa)
    torch.set_num_threads(os.cpu_count() or 1)
    sets = []
    for name in ("std_v010", "ens_v010", "std_v014", "ens_v014", "std_wex"):
        g, ck = load_golden(name)
        sd, meta = load_checkpoint(ck)
        x = oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], g.get("anns"), int(meta["coverage_thr"]))
        sets.append(("golden " + name, sd, x, (g["o1"], g["o2"], g["o3"])))
        if name in ("std_v014", "std_wex"):
            cand = synth.structured_pileups(6000, seed=901)
            xl = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, None, int(meta["coverage_thr"]))
            p = torch_port.prepare(sd)
            ref = [np.concatenate([np.asarray(torch_port.forward(p, torch.from_numpy(xl[i:i + 1000]))[j]) for i in range(0, len(xl), 1000)]) for j in range(3)]
            sets.append(("S2x6000 " + name, sd, xl, tuple(ref)))
    xt = torch.rand((32768, 26, 5, 32), device="cuda")
    out = []
    for mask in masks:
        os.environ["NSC_S8_MASK"] = str(mask)
        row = {"kappa": kappa, "mask": "0x%02x" % mask, "sets": {}}
        for name, sd, x, ref in sets:
            eng = Engine(sd, x.shape[1], precision=3)
            outs = [t.cpu().numpy() for t in eng.forward_f32(torch.from_numpy(x).cuda(), extras=True)]
            eng.close()
            row["sets"][name] = metrics(outs, *ref)
        eng = Engine(sets[0][1], 26, precision=3)
        for _ in range(3):
            eng.forward_f32(xt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            eng.forward_f32(xt)
        e1.record()
        torch.cuda.synchronize()
        eng.close()
        row["ms_per_32768"] = e0.elapsed_time(e1) / 10
        row["mm4_mean"] = float(np.mean([v["mm4"] for v in row["sets"].values()]))
        row["mm4_max"] = float(np.max([v["mm4"] for v in row["sets"].values()]))
        out.append(row)
        print(json.dumps(row), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "s8_mask_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
