"""Large-N parity of the CUDA path against the CPU port of the reference -- run on the GPU box.

    python tools/parity_large.py [--n 200000] [--modes 3,0,2] [--out gpurun_out/parity_large.json]

For C = 26 and C = 119, two candidate families of ``n`` candidates each:
  S2  structured pileups through bundled weights (v0.1.4 standalone / ensemble, tests/golden/*.pth)
  S3  the same pileups through a seeded randomly initialised network (near-tied logits: argmax is maximally sensitive)
and, per arithmetic mode (NSC_PREC_*), against ``oracle/torch_port.py`` on the box's host cores (fp32, the operators the
reference itself runs):
  argmax flips (type / length) out of n, max |d softmax|, max |d position|, max |d logit|, the share of 4-decimal
  probability entries that differ (call.py:97-100 rounds them; QUAL/SCORE in pred.vcf are functions of these), and the
  share of candidates with any differing 4-decimal probability.
Plus the VCF-line identity of the whole host stage: ``synth.genome_candidates`` (pileups + tags + FASTA that agree) ->
dictionaries from the GPU and from the oracle -> vcf.py -> pred.vcf / none.vcf lines, compared line by line.

The generators are chunked (20 000 candidates per seed) so that the float64 staging of the synthetic pileups stays
below 1 GB.  The oracle is the checker here, never the thing measured.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neusomatic_b200 import Engine, anns_to_255, synth, vcf                 # noqa: E402
from neusomatic_b200.call import PredTable, load_checkpoint, non_transformed_images   # noqa: E402
from neusomatic_b200.fasta import MemoryFasta                               # noqa: E402
from oracle import nsnet_oracle as oracle, torch_port                       # noqa: E402

CHUNK = 20000
MODE_NAMES = {0: "SPLIT16", 2: "FP32", 3: "SPLIT8"}


def chunks(n, seed0, ensemble, genome=False):
    for k, lo in enumerate(range(0, n, CHUNK)):
        m = min(CHUNK, n - lo)
        if genome:
            yield synth.genome_candidates(m, seed=seed0 + k, ensemble=ensemble)
        else:
            yield synth.structured_pileups(m, seed=seed0 + k, ensemble=ensemble)


def cpu_forward(p, x, batch=1000):
    outs = [[], [], []]
    for lo in range(0, len(x), batch):
        o = torch_port.forward(p, torch.from_numpy(x[lo:lo + batch]))
        for j in range(3):
            outs[j].append(np.asarray(o[j]))
    return [np.concatenate(o) for o in outs]


class Acc:
    def __init__(self):
        self.n = 0
        self.flip1 = self.flip3 = 0
        self.dprob = self.dpos = self.dlogit = 0.0
        self.mm_entries = self.entries = self.mm_cands = 0

    def add(self, g, r):
        o1, o2, o3, p1, p3, a1, a3 = g
        r1, r2, r3 = r
        self.n += len(r1)
        self.flip1 += int((a1 != r1.argmax(1)).sum())
        self.flip3 += int((a3 != r3.argmax(1)).sum())
        q1, q3 = oracle.softmax(r1), oracle.softmax(r3)
        self.dprob = max(self.dprob, float(np.abs(p1 - q1).max()), float(np.abs(p3 - q3).max()))
        self.dpos = max(self.dpos, float(np.abs(o2 - r2).max()))
        self.dlogit = max(self.dlogit, float(np.abs(o1 - r1).max()), float(np.abs(o3 - r3).max()))
        mm = np.concatenate([np.round(p1, 4) != np.round(q1, 4), np.round(p3, 4) != np.round(q3, 4)], 1)
        self.mm_entries += int(mm.sum())
        self.entries += mm.size
        self.mm_cands += int(mm.any(1).sum())

    def report(self):
        return {"n": self.n, "flips_type": self.flip1, "flips_len": self.flip3, "max_dprob": self.dprob, "max_dpos": self.dpos,
                "max_dlogit": self.dlogit, "mismatch_4dp_entries": self.mm_entries / max(self.entries, 1),
                "mismatch_4dp_candidates": self.mm_cands / max(self.n, 1)}


def run_family(name, sd, C, thr, n, seed0, modes, dev):
    ensemble = C == 119
    p = torch_port.prepare(sd)
    engines = {m: Engine(sd, C, precision=m) for m in modes}
    acc = {m: Acc() for m in modes}
    t_cpu = 0.0
    for cand in chunks(n, seed0, ensemble):
        x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns, thr)
        t0 = time.perf_counter()
        ref = cpu_forward(p, x)
        t_cpu += time.perf_counter() - t0
        a255 = anns_to_255(cand.anns)
        for m, eng in engines.items():
            g = eng.score_host_u8(cand.matrices, cand.meta, a255, float(thr), extras=True)
            acc[m].add([t.numpy() for t in g], ref)
    for eng in engines.values():
        eng.close()
    out = {"family": name, "C": C, "cpu_port_candidates_per_s": n / t_cpu, "modes": {MODE_NAMES[m]: acc[m].report() for m in modes}}
    print(json.dumps(out), flush=True)
    return out


def tables_from(outs, tags, matrices, center):
    o1, o2, o3, p1, p3, a1, a3 = outs
    ft, nt, tp = PredTable(), PredTable(), {}
    var = np.nonzero(a1 != 2)[0]
    non = np.nonzero(a1 == 2)[0]
    for tab, idx in ((ft, var), (nt, non)):
        tab.add([tags[i] for i in idx], a1[idx], o2[idx], a3[idx], p1[idx], p3[idx], o1[idx], o3[idx])
    img = non_transformed_images(matrices[var], center[var])
    for i, im in zip(var, img):
        tp[tags[i]] = im
    return ft, nt, tp


def vcf_identity_family(name, ck, n, seed0, modes):
    """pred.vcf / none.vcf lines from GPU dictionaries vs from the oracle's dictionaries, same host stage (vcf.py)."""
    sd, meta = load_checkpoint(ck)
    C = sd["conv1.weight"].shape[1]
    thr = int(meta["coverage_thr"])
    p = torch_port.prepare(sd)
    engines = {m: Engine(sd, C, precision=m) for m in modes}
    stat = {m: {"pred_lines": 0, "pred_identical": 0, "pred_same_record": 0, "none_lines": 0, "none_identical": 0,
                "records_only_in_one": 0} for m in modes}
    for cand, fasta, chroms in chunks(n, seed0, C == 119, genome=True):
        order = {c: i for i, c in enumerate(chroms)}
        fa = MemoryFasta(fasta)
        x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns, thr)
        r1, r2, r3 = cpu_forward(p, x)
        ref_outs = [r1, r2, r3, oracle.softmax(r1), oracle.softmax(r3), r1.argmax(1), r3.argmax(1)]

        def lines(outs):
            ft, nt, tp = tables_from(outs, cand.tags, cand.matrices, cand.center)
            recs = dict(vcf.pred_vcf_records(fa, ft, tp, chroms, oracle.VARTYPE_CLASSES, 1))
            none = dict(vcf.pred_vcf_records_none(nt, chroms))
            return (vcf.vcf_lines(vcf.get_vcf_records(recs), order, 0.7, 0.4),
                    vcf.vcf_lines(vcf.get_vcf_records(none), order, 0.7, 0.4))
        ref_pred, ref_none = lines(ref_outs)
        a255 = anns_to_255(cand.anns)
        for m, eng in engines.items():
            g = [t.numpy() for t in eng.score_host_u8(cand.matrices, cand.meta, a255, float(thr), extras=True)]
            got_pred, got_none = lines(g)
            s = stat[m]
            key = lambda l: tuple(l.split("\t")[:5])
            rp, gp = {key(l): l for l in ref_pred}, {key(l): l for l in got_pred}
            s["pred_lines"] += len(ref_pred)
            s["pred_identical"] += sum(1 for k, l in rp.items() if gp.get(k) == l)
            s["pred_same_record"] += sum(1 for k in rp if k in gp)
            s["records_only_in_one"] += len(set(rp) ^ set(gp))
            rn, gn = {key(l): l for l in ref_none}, {key(l): l for l in got_none}
            s["none_lines"] += len(ref_none)
            s["none_identical"] += sum(1 for k, l in rn.items() if gn.get(k) == l)
    for eng in engines.values():
        eng.close()
    out = {"family": name, "C": C, "n": n, "modes": {}}
    for m, s in stat.items():
        s["pred_identical_share"] = s["pred_identical"] / max(s["pred_lines"], 1)
        s["none_identical_share"] = s["none_identical"] / max(s["none_lines"], 1)
        out["modes"][MODE_NAMES[m]] = s
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200000)
    ap.add_argument("--n-vcf", type=int, default=100000)
    ap.add_argument("--modes", default="3,0,2")
    ap.add_argument("--families", default="S2_std,S3_std,S2_ens,S3_ens,VCF_std,VCF_wex")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_large.json"))
    args = ap.parse_args()
    modes = [int(m) for m in args.modes.split(",")]
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda", 0)
    G = os.path.join(ROOT, "tests", "golden")
    res = {"host_threads": torch.get_num_threads(), "families": []}
    fam = set(args.families.split(","))
    if "S2_std" in fam:
        sd, meta = load_checkpoint(os.path.join(G, "std_v014.pth"))
        res["families"].append(run_family("S2 structured pileups, v0.1.4 standalone weights", sd, 26, int(meta["coverage_thr"]), args.n, 50000, modes, dev))
    if "S3_std" in fam:
        res["families"].append(run_family("S3 structured pileups, random weights (seed 0)", synth.random_state_dict(26, seed=0), 26, 100, args.n, 60000, modes, dev))
    if "S2_ens" in fam:
        sd, meta = load_checkpoint(os.path.join(G, "ens_v014.pth"))
        res["families"].append(run_family("S2 structured pileups, v0.1.4 ensemble weights", sd, 119, int(meta["coverage_thr"]), args.n, 70000, modes, dev))
    if "S3_ens" in fam:
        res["families"].append(run_family("S3 structured pileups, random weights (seed 1), C=119", synth.random_state_dict(119, seed=1), 119, 100, args.n, 80000, modes, dev))
    if "VCF_std" in fam:
        res["families"].append(vcf_identity_family("VCF lines, genome candidates, v0.1.4 standalone weights", os.path.join(G, "std_v014.pth"), args.n_vcf, 90000, modes))
    if "VCF_wex" in fam:
        res["families"].append(vcf_identity_family("VCF lines, genome candidates, v0.1.0 WEX standalone weights", os.path.join(G, "std_wex.pth"), args.n_vcf, 95000, modes))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
