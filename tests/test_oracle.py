"""CPU: pin the numpy oracle against outputs of the unmodified reference (tests/golden)."""
import numpy as np

from oracle import nsnet_oracle as oracle


def _assemble(g, ck):
    return oracle.assemble_channels(
        g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], g.get("anns"),
        int(ck["coverage_thr"]), bool(ck.get("normalize_channels", False)))


def test_assemble_channels_bit_exact(golden):
    name, g, ck = golden
    x = _assemble(g, ck)
    assert x.dtype == np.float32 and x.shape[1] == (119 if "anns" in g else 26)
    # the reference's NeuSomaticDataset output for the first candidates: bit-exact
    np.testing.assert_array_equal(x[:4], g["x_first"])
    nt = oracle.non_transformed(g["matrices"][:8], g["center"][:8])
    np.testing.assert_array_equal(nt, g["non_transformed_first"])


def test_forward_matches_reference(golden):
    name, g, ck = golden
    p = oracle.normalize_state_dict(ck["state_dict"])
    (o1, o2, o3), internals = oracle.forward(p, _assemble(g, ck), return_internals=True)
    # fp32 summation-order noise only (reference = PyTorch CPU fp32)
    np.testing.assert_allclose(internals[0][:2], g["pool1_first"], atol=2e-5)
    np.testing.assert_allclose(internals[1][:8], g["res_first"], atol=1e-4)
    np.testing.assert_allclose(internals[3][:8], g["fc1_first"], atol=1e-4)
    np.testing.assert_allclose(o1, g["o1"], atol=2e-4)
    np.testing.assert_allclose(o2, g["o2"], atol=2e-4)
    np.testing.assert_allclose(o3, g["o3"], atol=2e-4)
    assert (o1.argmax(1) == g["o1"].argmax(1)).all()
    assert (o3.argmax(1) == g["o3"].argmax(1)).all()
    assert np.abs(oracle.softmax(o1) - oracle.softmax(g["o1"])).max() < 1e-4


def test_forward_fp64_noise_floor(golden):
    name, g, ck = golden
    p64 = oracle.normalize_state_dict(ck["state_dict"], np.float64)
    o1, o2, o3 = oracle.forward(p64, _assemble(g, ck)[:32])
    np.testing.assert_allclose(o1, g["o1"][:32], atol=3e-4)
    np.testing.assert_allclose(o2, g["o2"][:32], atol=3e-4)


def test_tsv_roundtrip_and_tag_parse(golden):
    name, g, ck = golden
    anns = g.get("anns")
    for i in (0, 5, 17):
        line = oracle.encode_tsv_line(i, str(g["tags"][i]), g["matrices"][i],
                                      None if anns is None else anns[i])
        tag, im, a = oracle.decode_tsv_line(line)
        assert tag == str(g["tags"][i])
        np.testing.assert_array_equal(im, g["matrices"][i])
        if anns is not None:
            np.testing.assert_array_equal(np.float32(a), anns[i])
        vt, center, length, tc, nc = oracle.parse_tag(tag)
        assert (center, tc, nc) == (g["center"][i], g["tumor_cov"][i], g["normal_cov"][i])
        assert vt in oracle.VARTYPE_CLASSES


def test_call_variants_contract(golden):
    name, g, ck = golden
    tags = [str(t) for t in g["tags"]]
    final, none = oracle.call_variants(g["o1"], g["o2"], g["o3"], tags)
    assert len(final) + len(none) == len(set(tags))
    for d, is_none in ((final, False), (none, True)):
        for k, e in d.items():
            assert (e[0] == "NONE") == is_none
            assert isinstance(e[3][0], np.float32) and len(e[3]) == 4 and len(e[6]) == 4
            assert e[1].shape == (1,)
            assert abs(sum(float(v) for v in e[3]) - 1.0) < 1e-3


def test_torch_port_matches_reference(golden):
    """oracle/torch_port.py (the timed CPU baseline) against the reference's own outputs."""
    import torch
    from oracle import torch_port
    name, g, ck = golden
    p = torch_port.prepare(ck["state_dict"])
    o1, o2, o3 = torch_port.forward(p, torch.from_numpy(_assemble(g, ck)))
    np.testing.assert_allclose(o1.numpy(), g["o1"], atol=1e-5)
    np.testing.assert_allclose(o2.numpy(), g["o2"], atol=1e-5)
    np.testing.assert_allclose(o3.numpy(), g["o3"], atol=1e-5)


def test_train_port_matches_reference_train_step():
    """oracle/train_port.py against two optimisation steps of the unmodified reference (train.py:436-441)."""
    import os
    import torch
    from conftest import GOLDEN
    from neusomatic_b200 import synth
    from oracle import train_port
    g = dict(np.load(os.path.join(GOLDEN, "train_std.npz")))
    p = {k: torch.from_numpy(v) for k, v in synth.random_state_dict(26, seed=11).items()}
    x = torch.from_numpy(oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], None, 100))
    labels, len_labels = torch.from_numpy(g["labels"]).long(), torch.from_numpy(g["len_labels"]).long()
    centers = torch.from_numpy(g["center"]).float()
    mom = {}
    for step in range(2):
        loss, parts, grads = train_port.grads(p, x, labels, centers, len_labels, torch.from_numpy(g["w_type"]),
                                              torch.from_numpy(g["w_len"]))
        np.testing.assert_allclose([loss.item()] + [t.item() for t in parts], g["loss%d" % step], rtol=2e-5)
        if step == 0:
            for k in g:
                if k.startswith("grad."):
                    np.testing.assert_allclose(grads[k[5:]].numpy(), g[k], rtol=1e-3, atol=max(1e-4 * float(np.abs(g[k]).max()), 1e-6))   # conv biases in front of a BN have ~0 gradient
        train_port.sgd_step(p, grads, mom, 0.01, 0.9)
    for k in g:
        if k.startswith("final."):
            np.testing.assert_allclose(p[k[6:]].numpy(), g[k], rtol=1e-3, atol=1e-4)


def test_split8_operand_arithmetic_meets_the_parity_bar():
    """The operand formats of NSC_PREC_SPLIT8 (fp16 main product + e4m3 corrections scaled by 2^15), emulated on the CPU
    with exact accumulation, stay inside the north-star tolerances on reference outputs; single-pass fp16 does not."""
    import torch
    from conftest import load_golden
    from oracle import split8_emulation as emu
    g, ck = load_golden("std_v014")
    p = {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}
    n = 48
    x = torch.from_numpy(oracle.assemble_channels(g["matrices"][:n], g["center"][:n], g["tumor_cov"][:n], g["normal_cov"][:n],
                                                  None, int(ck["coverage_thr"])))
    err = {}
    for mode in ("split8", "single16"):
        o1, o2, o3 = emu.forward(p, x, mode)
        dp = max(np.abs(oracle.softmax(o1) - oracle.softmax(g["o1"][:n])).max(), np.abs(oracle.softmax(o3) - oracle.softmax(g["o3"][:n])).max())
        err[mode] = (dp, np.abs(o2 - g["o2"][:n]).max())
        if mode == "split8":
            assert (o1.argmax(1) == g["o1"][:n].argmax(1)).all() and (o3.argmax(1) == g["o3"][:n].argmax(1)).all()
    assert err["split8"][0] <= 2e-3 / 4 and err["split8"][1] <= 1e-3 / 4, err      # 4x margin on both tolerances
    assert err["single16"][1] > err["split8"][1] * 4, err
