"""Small training steps for compute-sanitizer (run: compute-sanitizer --tool memcheck python tools/sanitize_train.py):
both arithmetic paths, standalone and ensemble channels, ragged batch sizes (13, 24), two steps each."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neusomatic_b200 import synth
from neusomatic_b200.train import Trainer
from oracle import nsnet_oracle as oracle
dev = "cuda"
for mode in ("tensor", "fp32"):
    os.environ.pop("NSC_TRAIN_FP32", None)
    if mode == "fp32":
        os.environ["NSC_TRAIN_FP32"] = "1"
    for C, B in ((26, 13), (119, 24)):
        cand = synth.structured_pileups(B, seed=5, ensemble=(C == 119))
        x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns if C == 119 else None, 100)
        labels = np.array([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], np.int32)
        lens = np.array([min(int(t.split(".")[6]), 3) for t in cand.tags], np.int32)
        w1 = torch.tensor([0.2, 0.21, 0.13, 0.16], device=dev); w3 = torch.tensor([0.12, 0.2, 0.23, 0.24], device=dev)
        tr = Trainer(synth.random_state_dict(C, seed=3), C, max_batch=B)
        for _ in range(2):
            losses = tr.step(torch.from_numpy(x).to(dev), torch.from_numpy(labels).to(dev), torch.from_numpy(cand.center).float().to(dev),
                             torch.from_numpy(lens).to(dev), w1, w3)
        torch.cuda.synchronize()
        print("%s C=%d B=%d ok: loss %.4f" % (mode, C, B, float(losses[0])), flush=True)
        tr.close()
