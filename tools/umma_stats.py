import os, sys, numpy as np, torch
os.environ["NSC_UMMA_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neusomatic_b200 import Engine, synth
from neusomatic_b200.call import load_checkpoint
sd, meta = load_checkpoint(os.path.join(ROOT, "tests/golden/std_v010.pth"))
eng = Engine(sd, 26, precision=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
x = torch.randn(8192, 26, 5, 32, device="cuda")
eng.forward_f32(x); torch.cuda.synchronize()
print("---- second pass ----", file=sys.stderr)
eng.forward_f32(x); torch.cuda.synchronize()
