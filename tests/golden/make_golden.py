#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What it does, per model (standalone v0.1.0, ensemble v0.1.0, standalone v0.1.4, ensemble v0.1.4, standalone WEX):
  1. generates seeded structured candidates (neusomatic_b200.synth.structured_pileups),
  2. writes them as a candidate TSV + .idx in the reference's on-disk format
     (generate_dataset.py:1569-1583),
  3. loads them back through the reference's own ``NeuSomaticDataset`` (is_test=True,
     matrix_transform as call.py:408,509-514) -- with the one-line numpy>=2 shim for
     ``extract_zlib`` (np.fromstring -> np.frombuffer, dataloader.py:38-39),
  4. runs the reference ``NeuSomaticNet`` (network.py) on CPU with the bundled checkpoint,
  5. stores compact inputs + reference outputs (+ a few assembled matrices / internal
     activations for layer-level debugging) as .npz, and re-saves the checkpoint dict in the
     reference's .pth layout (train.py:390-395) so the GPU box, which has no /root/reference,
     has weights to load.
"""
import os
import pickle
import sys
import tempfile
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/neusomatic"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "python"))

import dataloader as ref_dataloader            # noqa: E402  (reference)
from network import NeuSomaticNet              # noqa: E402  (reference)

from neusomatic_b200 import synth              # noqa: E402
from oracle import nsnet_oracle as oracle      # noqa: E402

ref_dataloader.extract_zlib = lambda b: np.frombuffer(
    zlib.decompress(b), np.uint8).reshape(5, 32, 23).copy()

MODELS = {
    "std_v010": ("NeuSomatic_v0.1.0_standalone_Dream3_70purity.pth", False, 192, 2024),
    "ens_v010": ("NeuSomatic_v0.1.0_ensemble_Dream3_70purity.pth", True, 96, 2025),
    "std_v014": ("NeuSomatic_v0.1.4_standalone_SEQC-WGS-Spike.pth", False, 192, 2026),
    "ens_v014": ("NeuSomatic_v0.1.4_ensemble_SEQC-WGS-GT50-SpikeWGS10.pth", True, 96, 2027),
    "std_wex": ("NeuSomatic_v0.1.0_standalone_WEX_100purity.pth", False, 128, 2028),
}


def main():
    torch.set_num_threads(1)
    for name, (fn, ensemble, n, seed) in MODELS.items():
        ck = torch.load(os.path.join(REF, "models", fn), map_location="cpu", weights_only=True)
        torch.save(ck, os.path.join(HERE, name + ".pth"))
        sd = {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}
        C = 119 if ensemble else 26
        net = NeuSomaticNet(C)
        net.load_state_dict(sd)
        net.eval()
        cand = synth.structured_pileups(n, seed=seed, ensemble=ensemble)
        with tempfile.TemporaryDirectory() as td:
            tsv = os.path.join(td, "candidates_0.tsv")
            offs = []
            with open(tsv, "w") as f:
                for i in range(n):
                    offs.append(f.tell())
                    f.write(oracle.encode_tsv_line(
                        i, cand.tags[i], cand.matrices[i],
                        None if cand.anns is None else cand.anns[i]))
                offs.append(f.tell())
            pickle.dump(offs, open(tsv + ".idx", "wb"))
            ds = ref_dataloader.NeuSomaticDataset(
                roots=[tsv], max_load_candidates=100000,
                transform=ref_dataloader.matrix_transform((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
                is_test=True, num_threads=1, coverage_thr=ck["coverage_thr"],
                normalize_channels=ck.get("normalize_channels", False))
            xs, nts, paths = [], [], []
            for i in range(n):
                (matrix, _label, _var_pos, _vl, nt), (path, _) = ds[i]
                xs.append(matrix.contiguous().numpy())
                nts.append(nt)
                paths.append(path)
        x = np.ascontiguousarray(np.stack(xs), dtype=np.float32)   # torch.stack in default_collate
        assert paths == cand.tags
        with torch.no_grad():
            (o1, o2, o3), internals = net(torch.from_numpy(x))
        out = dict(
            matrices=cand.matrices, center=cand.center, tumor_cov=cand.tumor_cov,
            normal_cov=cand.normal_cov, tags=np.array(cand.tags),
            coverage_thr=np.int64(ck["coverage_thr"]),
            o1=o1.numpy(), o2=o2.numpy(), o3=o3.numpy(),
            x_first=x[:4], non_transformed_first=np.stack(nts[:8]),
            pool1_first=internals[0][:2].numpy(), res_first=internals[1][:8].numpy(),
            fc1_first=internals[3][:8].numpy())
        if cand.anns is not None:
            out["anns"] = cand.anns
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        p = torch.softmax(o1, 1)
        print(name, "n=%d" % n, "classes", np.bincount(o1.argmax(1).numpy(), minlength=4),
              "frac maxprob<0.999: %.3f" % float((p.max(1).values < 0.999).float().mean()))


if __name__ == "__main__":
    main()
