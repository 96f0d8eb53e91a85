"""2+ GPU check of the data-parallel train step (run under torchrun on the GPU box):
every rank runs nsc_train_forward_backward on its own shard with the loss normalised by the GLOBAL class-weight sums /
batch size (3 floats all-reduced up front: the reference's nn.DataParallel computes one loss over the gathered batch,
train.py:436-441), the flat gradient vector is summed with ONE NCCL all-reduce, nsc_sgd_momentum_update applies it.
Checks: (1) all ranks end with identical parameters, (2) they equal a single-process update with the sum of the per-shard
gradients of the global loss, (3) the per-rank losses sum to the global loss; prints step throughput."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neusomatic_b200 import synth, _lib
from neusomatic_b200.train import Trainer
from oracle import nsnet_oracle as oracle

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sd = synth.random_state_dict(26, seed=11)
dev = torch.device("cuda", lr_)

def shard(r):
    cand = synth.structured_pileups(B, seed=100 + r)
    x = torch.from_numpy(oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, None, 100)).to(dev)
    vt = torch.tensor([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], dtype=torch.int32, device=dev)
    ln = torch.tensor([min(int(t.split(".")[6]), 3) for t in cand.tags], dtype=torch.int32, device=dev)
    return x, vt, torch.from_numpy(cand.center).float().to(dev), ln

w1 = torch.tensor([0.2, 0.21, 0.13, 0.16], device=dev); w3 = torch.tensor([0.12, 0.2, 0.23, 0.24], device=dev)
tr = Trainer(sd, 26, device=lr_, max_batch=B)
x, vt, ce, ln = shard(rank)
losses = tr.step(x, vt, ce, ln, w1, w3).clone()
torch.cuda.synchronize()
loss_sum = losses.clone(); dist.all_reduce(loss_sum)
# (1) identical parameters on every rank
ref = tr.params.clone(); dist.broadcast(ref, 0)
assert torch.equal(ref, tr.params), "ranks diverged"
# (2) equals the sum of the per-shard gradients of the global loss (recomputed locally on rank 0)
if rank == 0:
    shards = [shard(r) for r in range(world)]
    norm = torch.stack([sum(w1[s[1].long()].sum() for s in shards), sum(w3[s[3].long()].sum() for s in shards),
                        torch.tensor(float(world * B), device=dev)]).float()
    gsum, lsum = None, 0.0
    for r in range(world):
        t2 = Trainer(sd, 26, device=lr_, max_batch=B)
        _lib.check(t2.lib.nsc_trainer_set_loss_norm(t2.h, norm.data_ptr()))
        lsum = lsum + t2.forward_backward(*shards[r], w1, w3).clone()
        gsum = t2.grads.clone() if gsum is None else gsum + t2.grads
        t2.close()
    t3 = Trainer(sd, 26, device=lr_, max_batch=B)
    expect = t3.params - 0.01 * gsum
    err = (expect - tr.params).abs().max().item()
    print("max |param - expected| after the data-parallel step: %.3e; global loss %.6f (sum of rank losses %.6f)" % (
        err, float(lsum[0]), float(loss_sum[0])))
    assert err < 1e-6 and abs(float(lsum[0]) - float(loss_sum[0])) < 1e-5 * abs(float(lsum[0]))
# throughput
for _ in range(3):
    tr.step(x, vt, ce, ln, w1, w3)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
K = 10
for _ in range(K):
    tr.step(x, vt, ce, ln, w1, w3)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    print("train step: world=%d batch/GPU=%d  %.2f ms/step  %.0f samples/s (whole job), grad vector %d floats" % (
        world, B, 1000 * dt / K, world * B * K / dt, tr.n))
dist.destroy_process_group()
