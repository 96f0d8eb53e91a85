"""Sweep of the round-toward-zero compensation strength kappa (csrc/nsc_umma.cu: rz_compensation) -- run on the GPU box.

    python tools/rz_kappa_sweep.py [kappa ...]

For every kappa and tensor mode the five golden sets (reference outputs of the unmodified network.py, tests/golden) and a
larger S2 set per bundled checkpoint (against oracle/torch_port.py on the host cores) are scored; printed per set: the share
of 4-decimal probability entries that differ from the reference's, max |d softmax|, max |d logit|, max |d position| and the
mean signed error of the largest logit (the residual bias)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden                               # noqa: E402
from neusomatic_b200 import Engine, synth                     # noqa: E402
from neusomatic_b200.call import load_checkpoint              # noqa: E402
from oracle import nsnet_oracle as oracle, torch_port         # noqa: E402


def metrics(outs, r1, r2, r3):
    o1, o2, o3, p1, p3 = outs[:5]
    q1, q3 = oracle.softmax(r1), oracle.softmax(r3)
    mm = np.concatenate([np.round(p1, 4) != np.round(q1, 4), np.round(p3, 4) != np.round(q3, 4)], 1)
    top = r1.argmax(1)
    return {"mm4": float(mm.mean()), "dprob": float(max(np.abs(p1 - q1).max(), np.abs(p3 - q3).max())),
            "dlogit": float(max(np.abs(o1 - r1).max(), np.abs(o3 - r3).max())), "dpos": float(np.abs(o2 - r2).max()),
            "bias_top": float((o1[np.arange(len(top)), top] - r1[np.arange(len(top)), top]).mean()),
            "flips": int((o1.argmax(1) != r1.argmax(1)).sum() + (o3.argmax(1) != r3.argmax(1)).sum())}


def main():
    kappas = [float(a) for a in sys.argv[1:]] or [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0]
    torch.set_num_threads(os.cpu_count() or 1)
    sets = []
    for name in ("std_v010", "ens_v010", "std_v014", "ens_v014", "std_wex"):
        g, ck = load_golden(name)
        sd, meta = load_checkpoint(ck)
        x = oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], g.get("anns"), int(meta["coverage_thr"]))
        sets.append(("golden " + name, sd, x, (g["o1"], g["o2"], g["o3"])))
        if name in ("std_v014", "std_wex", "ens_v014"):
            cand = synth.structured_pileups(6000, seed=901, ensemble="anns" in g)
            xl = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns, int(meta["coverage_thr"]))
            p = torch_port.prepare(sd)
            ref = [np.concatenate([np.asarray(torch_port.forward(p, torch.from_numpy(xl[i:i + 1000]))[j]) for i in range(0, len(xl), 1000)]) for j in range(3)]
            sets.append(("S2x6000 " + name, sd, xl, tuple(ref)))
    out = []
    for kappa in kappas:
        os.environ["NSC_RZ_KAPPA"] = repr(kappa)
        for mode in (0, 3):
            row = {"kappa": kappa, "mode": {0: "SPLIT16", 3: "SPLIT8"}[mode], "sets": {}}
            for name, sd, x, ref in sets:
                eng = Engine(sd, x.shape[1], precision=mode)
                outs = [t.cpu().numpy() for t in eng.forward_f32(torch.from_numpy(x).cuda(), extras=True)]
                eng.close()
                row["sets"][name] = metrics(outs, *ref)
            row["mm4_mean"] = float(np.mean([v["mm4"] for v in row["sets"].values()]))
            row["mm4_max"] = float(np.max([v["mm4"] for v in row["sets"].values()]))
            out.append(row)
            print(json.dumps(row), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rz_kappa_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
