"""Host-side logic of the drop-in layers that needs no GPU: names, packing, parameter compatibility, synthetic configs."""
import pytest
import torch

import pna_b200
from pna_b200 import _lib, synth
from pna_b200.aggregate import output_width
from conftest import load_golden


def test_code_packing_follows_ctor_order():
    n, codes = _lib.pack_codes(["mean", "max", "min", "std"], _lib.AGGR_CODES, "aggregator")
    assert n == 4 and [(codes >> (4 * i)) & 15 for i in range(4)] == [1, 3, 2, 5]
    n, codes = _lib.pack_codes("identity amplification attenuation", _lib.SCALER_CODES, "scaler")   # DGL string form
    assert n == 3 and [(codes >> (4 * i)) & 15 for i in range(3)] == [0, 1, 2]
    with pytest.raises(KeyError):
        _lib.pack_codes(["moment3"], _lib.AGGR_CODES, "aggregator")
    assert output_width(75, 4, 3, False) == 900 and output_width(15, 4, 3, True) == 195


def test_avg_deg_matches_reference_ctor():
    for name in ("pyg_simple_f16", "pyg_conv_t4_div"):
        g = load_golden(name)
        mine = pna_b200.avg_deg_from_histogram(g["deg"])
        assert mine["log"] == g["avg_deg"]["log"] and mine["lin"] == g["avg_deg"]["lin"]


@pytest.mark.parametrize("name", ["pyg_simple_f16", "pyg_simple_allops"])
def test_reference_state_dict_loads_into_simple_layer(name):
    g = load_golden(name)
    f = g["x"].size(1)
    lay = pna_b200.PNAConvSimple(f, f, g["aggregators"], g["scalers"], g["deg"], post_layers=g["post_layers"])
    assert set(lay.state_dict()) == set(g["state_dict"])
    lay.load_state_dict(g["state_dict"], strict=True)


@pytest.mark.parametrize("name", ["pyg_conv_t1", "pyg_conv_t4_div", "pyg_conv_t5_rep", "pyg_conv_edge", "pyg_conv_pre2"])
def test_reference_state_dict_loads_into_conv_layer(name):
    g = load_golden(name)
    c = g["ctor"]
    lay = pna_b200.PNAConv(c["in_channels"], c["out_channels"], g["aggregators"], g["scalers"], g["deg"], edge_dim=c["edge_dim"],
                           towers=c["towers"], pre_layers=c["pre_layers"], post_layers=c["post_layers"],
                           divide_input=c["divide_input"])
    assert set(lay.state_dict()) == set(g["state_dict"])
    for k, v in lay.state_dict().items():
        assert v.shape == g["state_dict"][k].shape, k
    lay.load_state_dict(g["state_dict"], strict=True)


def test_affine_message_decomposition_equals_per_edge_linear():
    """m_e = pre_nn([x_i || x_j]) == U[i] + V[j] with U = x W_i^T, V = x W_j^T + b (pna.py:94,147-149)."""
    g = load_golden("pyg_conv_t4_div")
    c = g["ctor"]
    lay = pna_b200.PNAConv(c["in_channels"], c["out_channels"], g["aggregators"], g["scalers"], g["deg"], towers=c["towers"],
                           post_layers=c["post_layers"], divide_input=True)
    lay.load_state_dict(g["state_dict"])
    x, ei = g["x"], g["edge_index"]
    with torch.no_grad():
        U, V = lay._affine_terms(x, lay.F_in)
        xt = x.view(-1, lay.towers, lay.F_in)
        h = torch.cat([xt[ei[1]], xt[ei[0]]], -1)
        msg = torch.stack([nn(h[:, t]) for t, nn in enumerate(lay.pre_nns)], 1).reshape(ei.size(1), -1)
    torch.testing.assert_close(U[ei[1]] + V[ei[0]], msg, rtol=1e-5, atol=1e-5)


def test_synthetic_configs_have_the_stated_shapes():
    ei, x = synth.arxiv_like(n_nodes=5000, n_edges=40000, n_feat=8)
    assert ei.shape == (2, 40000) and x.shape == (5000, 8) and int(ei.max()) < 5000
    deg = torch.bincount(ei[1], minlength=5000)
    assert int(deg.max()) > 200 and int((deg == 0).sum()) > 50          # skewed: hubs and isolated rows
    ei, x, batch = synth.zinc_like(n_graphs=50, n_feat=75)
    assert x.size(1) == 75 and batch.numel() == x.size(0)
    assert bool((batch[ei[0]] == batch[ei[1]]).all())                     # no edge crosses molecules
    assert torch.equal(torch.sort(ei[0] * x.size(0) + ei[1]).values, torch.sort(ei[1] * x.size(0) + ei[0]).values)  # symmetric
    ei, x = synth.superpixel_like(n_graphs=20, n_feat=4)
    assert ei.size(1) == 20 * 70 * 8 and bool(((ei[0] // 70) == (ei[1] // 70)).all())
    ei, x = synth.powerlaw(n_nodes=2000, n_edges=20000, n_feat=4)
    assert int(torch.bincount(ei[1], minlength=2000).max()) > 500
    ei, x = synth.multitask_like(n_graphs=3, nodes_per_graph=100)
    assert x.shape == (300, 16)
    b = synth.algorithmic_bytes(169343, 1166243, 128, 4, 1536)
    assert abs(b["b_min"] / 1e9 - 1.13) < 0.01 and abs(b["b_gather"] / 1e9 - 1.64) < 0.01


def test_padding_algebra_is_exact_on_the_oracle():
    """Zero-padded feature blocks + zero weight columns reproduce the unpadded layer output exactly (pna_b200/padding.py)."""
    from oracle import pna_oracle as O
    from pna_b200 import padding as pad
    torch.manual_seed(0)
    n, e, f = 60, 300, 15
    fp = pad.padded_width(f, torch.float32)
    assert fp == 16 and pad.padded_width(75, torch.bfloat16) == 80 and pad.padded_width(128, torch.float32) == 128
    x, ei = torch.randn(n, f), torch.randint(0, n, (2, e))
    A, S = ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"]
    avg = O.avg_deg_from_histogram(torch.bincount(torch.bincount(ei[1], minlength=n)))
    W, b = torch.randn(7, 12 * f), torch.randn(7)
    ref = torch.nn.functional.linear(O.simple_propagate(x, ei, A, S, avg), W, b)
    got = torch.nn.functional.linear(O.simple_propagate(pad.pad_cols(x, fp), ei, A, S, avg),
                                     pad.expand_weight_cols(W, 12, f, fp), b)
    assert torch.equal(ref, got)
    # tower blocks
    xt = torch.randn(n, 3 * f)
    p = pad.pad_blocks(xt, 3, f, fp)
    assert p.shape == (n, 3 * fp) and torch.equal(p.view(n, 3, fp)[:, :, :f], xt.view(n, 3, f)) and float(p.view(n, 3, fp)[:, :, f:].abs().sum()) == 0
    Wr = pad.expand_weight_rows(torch.randn(f, 9), f, fp)
    assert Wr.shape == (fp, 9) and float(Wr[f:].abs().sum()) == 0


def test_cache_keys_accept_inference_mode_tensors():
    """Tensors created under torch.inference_mode() do not track a version counter; the cache keys must not read it."""
    from pna_b200.csr import tensor_version
    with torch.inference_mode():
        t = torch.arange(6).view(2, 3)
    assert t.is_inference() and tensor_version(t) is None
    u = torch.arange(6)
    v0 = tensor_version(u)
    u.add_(1)
    assert tensor_version(u) == v0 + 1


@pytest.mark.parametrize("cin,cout,towers,divide", [(75, 75, 5, True), (15, 20, 5, False), (16, 16, 4, True), (32, 32, 1, False)])
def test_conv_tensor_core_weight_pack_is_the_reference_algebra(cin, cout, towers, divide):
    """PNAConv._tensor_core_pack (pna_b200/pyg.py): the zero-padded U|V weight, the BLOCK-DIAGONAL tower weight and the padded
    final Linear reproduce pna.py:131-135 when applied with plain matmuls on CPU (the GPU test runs them through
    pna_linear_fwd)."""
    import torch.nn.functional as Fn
    from pna_b200 import PNAConv, padding as pad
    torch.manual_seed(cin + towers)
    lay = PNAConv(cin, cout, ["mean", "max", "min", "std"], ["identity", "amplification", "attenuation"], torch.tensor([0, 4, 3, 1]),
                  towers=towers, divide_input=divide)
    T, Fi, Fo = towers, lay.F_in, lay.F_out
    Fp = pad.padded_width(Fi, torch.float32)
    with torch.no_grad():
        tc = lay._tensor_core_pack(Fp)
        x = torch.randn(23, cin)
        uv = pad.pad_cols(x, tc["K1"]) @ tc["w1"].t() + tc["b1"]
        U, V = lay._affine_terms(x, Fp)
        torch.testing.assert_close(uv[:, : T * Fp], U, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(uv[:, T * Fp: 2 * T * Fp], V, rtol=1e-6, atol=1e-6)
        assert uv[:, 2 * T * Fp:].abs().sum() == 0
        W = 13 * Fp
        out = torch.randn(23, T, W)                                   # stands for cat([x_t, aggregate_t]) at the padded width
        buf = torch.zeros(23, tc["K2"]); buf[:, : T * W] = out.reshape(23, T * W)
        h = buf @ tc["w2"].t() + tc["b2"]
        want_h = torch.cat([Fn.linear(out[:, t], pad.expand_weight_cols(nn[0].weight, 13, Fi, Fp), nn[0].bias)
                            for t, nn in enumerate(lay.post_nns)], 1)
        torch.testing.assert_close(h[:, : T * Fo], want_h, rtol=1e-5, atol=1e-5)
        assert h[:, T * Fo:].abs().sum() == 0
        y = (h @ tc["w3"].t() + tc["b3"])[:, :cout]
        torch.testing.assert_close(y, lay.lin(want_h), rtol=1e-5, atol=1e-5)
        # in-place parameter update -> new pack
        lay.lin.bias.add_(1.0)
        assert lay._tensor_core_pack(Fp) is not tc
