"""CPU, world_size 2 over gloo: the multi-GPU scoring path (call.call_sharded) shards candidates across
ranks without a data-path collective and rank 0 ends up with exactly the single-process dictionaries.
The per-rank scorer here is the oracle (no GPU in this container); on the GPU box it is CandidateScorer.call."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import nsnet_oracle as oracle
from neusomatic_b200 import synth
from neusomatic_b200.call import build_pred_dicts, call_sharded, shard_ranges


def _oracle_score_fn(ck):
    p = oracle.normalize_state_dict(ck["state_dict"])

    def fn(cand):
        if len(cand) == 0:
            return {}, {}
        x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns,
                                     int(ck["coverage_thr"]))
        o1, o2, o3 = oracle.forward(p, x)
        return build_pred_dicts(o1, o2, o3, oracle.softmax(o1), oracle.softmax(o3), o1.argmax(1), o3.argmax(1), cand.tags)
    return fn


def _cands(g, n):
    return synth.Candidates(g["matrices"][:n], g["center"][:n], g["tumor_cov"][:n], g["normal_cov"][:n], None,
                            [str(t) for t in g["tags"][:n]])


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, ck = load_golden("std_v010")
    f, nn_ = call_sharded(_cands(g, n), _oracle_score_fn(ck))
    if rank == 0:
        q.put((sorted(f), sorted(nn_), {k: [float(x) for x in v[3]] for k, v in {**f, **nn_}.items()}))
    else:
        assert f is None and nn_ is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [41, 64])
def test_two_rank_sharded_call_equals_single_process(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    keys_f, keys_n, probs = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g, ck = load_golden("std_v010")
    f1, n1 = _oracle_score_fn(ck)(_cands(g, n))
    assert keys_f == sorted(f1) and keys_n == sorted(n1)
    for k, v in {**f1, **n1}.items():
        assert [float(x) for x in v[3]] == probs[k]
    assert sum(b - a for a, b in shard_ranges(n, 2)) == n


def _tsv_worker(rank, world, port, path, q):
    """One rank of a sharded real-file call: its range of the TSV through run_tsv_pipeline; the
    scorer is the CPU stand-in of tests/test_call_host.py (on the GPU box it is CandidateScorer)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_call_host import _OracleScorer
    from neusomatic_b200.call import PredTable, run_tsv_pipeline
    from neusomatic_b200.tsv import CandidateTSV
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, ck = load_golden("std_v014")
    t = CandidateTSV(path)
    lo, hi = shard_ranges(len(t), world)[rank]
    t.close()
    fp, npd, tp = PredTable(), PredTable(), {}
    run_tsv_pipeline([_OracleScorer(ck, 26)], [(path, lo, hi - lo)], 32, 1, fp, npd, tp)
    mine = {"final": list(fp), "none": list(npd), "probs": {k: [float(x) for x in v[3]] for k, v in {**dict(fp), **dict(npd)}.items()},
            "images": sorted(tp)}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)          # control plane only: tags and a few floats, no data-path collective
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_tsv_call_equals_single_process(tmp_path):
    from test_call_host import _OracleScorer
    from neusomatic_b200.call import PredTable, run_tsv_pipeline
    n = 150
    cand = synth.structured_pileups(n, seed=17)
    path = str(tmp_path / "candidates.tsv")
    with open(path, "w") as f:
        for i in range(n):
            f.write(oracle.encode_tsv_line(i, cand.tags[i], cand.matrices[i]))
    _, ck = load_golden("std_v014")
    fp, npd, tp = PredTable(), PredTable(), {}
    run_tsv_pipeline([_OracleScorer(ck, 26)], [(path, 0, None)], 32, 1, fp, npd, tp)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tsv_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # rank order = candidate order: the concatenation of the ranks' tables is the single-process table
    assert parts[0]["final"] + parts[1]["final"] == list(fp) and parts[0]["none"] + parts[1]["none"] == list(npd)
    assert sorted(parts[0]["images"] + parts[1]["images"]) == sorted(tp) and len(fp) + len(npd) == n
    probs = {**parts[0]["probs"], **parts[1]["probs"]}
    for k, v in {**dict(fp), **dict(npd)}.items():          # (PyTorch's CPU kernels are not bit-stable across batch shapes: one unit
        assert np.abs(np.asarray(v[3], np.float64) - np.asarray(probs[k])).max() <= 1.01e-4      # of the 4th decimal at most)
