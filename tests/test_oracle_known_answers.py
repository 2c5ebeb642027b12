"""CPU tests pinning the oracle: hand-derived known-answer vectors, the ATen
properties the CUDA kernels rely on (F9, F10), the Philox known-answer vectors, and the golden
fixtures recorded from the reference's own nets / train_epoch (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import graph_build_np as G
from oracle import pyg_restated as O

PATH = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]])  # reference fixture: tests/test_masks_and_metrics.py:12
bits = lambda t: np.asarray(t, dtype=np.float32).view(np.uint32).tolist()


def test_symmetrise_known_answer():
    s = G.symmetrize(PATH.numpy())
    assert s.tolist() == [[0, 1, 2, 3, 1, 2, 3, 4], [1, 2, 3, 4, 0, 1, 2, 3]]
    assert np.array_equal(s, torch.cat([PATH, PATH.flip(0)], dim=1).numpy())  # train_gnn.py:321-324


def test_gcn_norm_known_answers_raw_path():
    ei2, w = O.gcn_norm(PATH, 5)
    assert ei2.tolist() == [[0, 1, 2, 3, 0, 1, 2, 3, 4], [1, 2, 3, 4, 0, 1, 2, 3, 4]]
    assert bits(w.numpy()) == [0x3F3504F3, 0x3EFFFFFF, 0x3EFFFFFF, 0x3EFFFFFF, 0x3F800000] + [0x3EFFFFFF] * 4
    ei2n, wn, dis = G.gcn_norm(PATH.numpy(), 5)
    assert np.array_equal(ei2n, ei2.numpy()) and bits(wn) == bits(w.numpy())
    assert bits(dis) == [0x3F800000] + [0x3F3504F3] * 4


def test_gcn_norm_known_answers_symmetrised_path():
    ei = torch.cat([PATH, PATH.flip(0)], dim=1)
    _, w = O.gcn_norm(ei, 5)
    assert bits(w.numpy()[:8]) == [0x3ED105EB, 0x3EAAAAAA, 0x3EAAAAAA, 0x3ED105EB] * 2
    assert bits(w.numpy()[8:]) == [0x3EFFFFFF, 0x3EAAAAAA, 0x3EAAAAAA, 0x3EAAAAAA, 0x3EFFFFFF]


def test_self_loop_rewrite_adversarial():
    ei = torch.tensor([[2, 0, 2, 0, 1], [2, 1, 2, 1, 0]])  # (2,2) twice, duplicate (0,1), reciprocal
    ei2 = O.add_remaining_self_loops(ei, 4)
    assert ei2.tolist() == [[0, 0, 1, 0, 1, 2, 3], [1, 1, 0, 0, 1, 2, 3]]
    assert np.array_equal(G.add_remaining_self_loops(ei.numpy(), 4), ei2.numpy())
    m = O.scatter_mean(torch.eye(4).index_select(0, ei[0]), ei[1], 4)  # SAGE counts both loops / dups
    assert m[2].tolist() == [0, 0, 1, 0] and m[1].tolist() == [1, 0, 0, 0] and m[3].tolist() == [0] * 4


def test_sage_mean_known_answers():
    x = torch.eye(5)
    m = O.SAGEConv(5, 3).aggregate(x, PATH)
    want = torch.zeros(5, 5)
    for i in range(1, 5):
        want[i, i - 1] = 1
    assert torch.equal(m, want)
    m = O.SAGEConv(5, 3).aggregate(x, torch.cat([PATH, PATH.flip(0)], 1))
    want = torch.zeros(5, 5)
    want[0, 1] = want[4, 3] = 1
    for i in (1, 2, 3):
        want[i, i - 1] = want[i, i + 1] = 0.5
    assert torch.equal(m, want)


def test_gat_zero_attention_is_mean_over_neighbours_and_self():
    conv = O.GATConv(5, 5, heads=1)
    with torch.no_grad():
        conv.lin.weight.copy_(torch.eye(5))
        conv.att_src.zero_()
        conv.att_dst.zero_()
    out = conv(torch.eye(5), PATH)
    want = torch.eye(5)
    for i in range(1, 5):
        want[i] = 0.5 * (torch.eye(5)[i] + torch.eye(5)[i - 1])
    assert torch.allclose(out, want, atol=1e-7)


def test_gat_backward_formulas_match_autograd_fp64():
    """The closed-form backward the CUDA kernels implement vs autograd, in fp64."""
    torch.manual_seed(0)
    n, H, C, fin = 30, 3, 4, 6
    ei = torch.randint(0, n, (2, 90))
    for concat in (True, False):
        conv = O.GATConv(fin, C, heads=H, concat=concat).double()
        x = torch.randn(n, fin, dtype=torch.float64, requires_grad=True)
        out = conv(x, ei)
        dout = torch.randn_like(out)
        out.backward(dout)
        with torch.no_grad():
            xs = (x @ conv.lin.weight.T).view(n, H, C)
            a_s, a_d = (xs * conv.att_src).sum(-1), (xs * conv.att_dst).sum(-1)
            ei2 = O.add_remaining_self_loops(ei, n)
            src, dst = ei2
            pre = a_s[src] + a_d[dst]
            e = torch.nn.functional.leaky_relu(pre, 0.2)
            p = (e - O.scatter_max(e, dst, n)[dst]).exp()
            alpha = p / (O.scatter_sum(p, dst, n) + 1e-16)[dst]
            do = dout.view(n, H, C) if concat else (dout / H).unsqueeze(1).expand(n, H, C)
            g = (do[dst] * xs[src]).sum(-1)
            s = O.scatter_sum(alpha * g, dst, n)
            dpre = alpha * (g - s[dst]) * torch.where(pre > 0, 1.0, 0.2)
            da_s, da_d = O.scatter_sum(dpre, src, n), O.scatter_sum(dpre, dst, n)
            dxs = (O.scatter_sum(alpha.unsqueeze(-1) * do[dst], src, n) + da_s.unsqueeze(-1) * conv.att_src
                   + da_d.unsqueeze(-1) * conv.att_dst)
            dW = dxs.view(n, H * C).T @ x
            dx = dxs.view(n, H * C) @ conv.lin.weight
            assert torch.allclose(dx, x.grad, atol=1e-12)
            assert torch.allclose(dW, conv.lin.weight.grad, atol=1e-12)
            assert torch.allclose((da_s.unsqueeze(-1) * xs).sum(0), conv.att_src.grad[0], atol=1e-12)
            assert torch.allclose((da_d.unsqueeze(-1) * xs).sum(0), conv.att_dst.grad[0], atol=1e-12)


def test_F9_scatter_add_is_sequential_in_edge_order():
    g = torch.Generator().manual_seed(0)
    n, e, f = 200, 5000, 8
    idx = torch.randint(0, n, (e,), generator=g)
    src = torch.randn(e, f, generator=g) * torch.exp(3 * torch.randn(e, 1, generator=g))
    got = O.scatter_sum(src, idx, n).numpy()
    want = np.zeros((n, f), dtype=np.float32)
    s, ix = src.numpy(), idx.numpy()
    for k in range(e):
        want[ix[k]] = want[ix[k]] + s[k]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_F10_pow_minus_half_is_one_over_sqrt():
    d = torch.arange(1, 20001, dtype=torch.float32)
    a = d.clone().pow_(-0.5).numpy()
    b = (np.float32(1) / np.sqrt(d.numpy())).astype(np.float32)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_sorted_views_reproduce_scatter_order():
    g = torch.Generator().manual_seed(5)
    n, e, f = 50, 400, 3
    ei = torch.randint(0, n, (2, e), generator=g)
    x = torch.randn(n, f, generator=g) * 100
    ptr, col, eid = G.sorted_view(ei.numpy(), n, 1)
    want = O.scatter_sum(x.index_select(0, ei[0]), ei[1], n).numpy()
    got = np.zeros((n, f), dtype=np.float32)
    for i in range(n):
        for p in range(ptr[i], ptr[i + 1]):
            got[i] = got[i] + x.numpy()[col[p]]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(ei.numpy()[0][eid], col) and np.all(np.diff(eid.reshape(-1)[ptr[3]:ptr[4]]) > 0)


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = G.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want


def test_dropout_mask_statistics_and_shard_invariance():
    m = G.dropout_keep_mask(42, 0, 0, 2000, 64, 0.2)
    assert abs(m.mean() - 0.8) < 0.01
    part = G.dropout_keep_mask(42, 0, 700, 300, 64, 0.2)
    assert np.array_equal(part, m[700:1000])
    assert not np.array_equal(G.dropout_keep_mask(42, 1, 0, 2000, 64, 0.2), m)
    assert G.dropout_threshold(0.5) == 2**15 and G.dropout_threshold(0.0) == 0


# ---- golden fixtures recorded from the reference's own nets / train_epoch ---------------------
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_nets.pt")


@pytest.mark.parametrize("name", ["gcn", "sage", "rec_k8", "gat"])
def test_oracle_reproduces_reference_fixtures(name):
    fx = torch.load(GOLD, weights_only=False)[name]
    cfg = fx["cfg"]
    torch.manual_seed(0)
    model = O.build_model(cfg["arch"], fx["x"].size(1), cfg)
    model.load_state_dict(fx["state0"])
    _, logits0 = O.eval_probs(model, fx["x"], fx["edge_index_used"], fx["timestep"])
    assert torch.allclose(logits0, fx["logits_eval0"], rtol=0, atol=1e-6)
    assert torch.equal(O.class_weight(fx["y"][fx["train_mask"]]), fx["class_weight"])
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"])
    for want in fx["losses"]:
        loss, _ = O.train_step(model, fx["x"], fx["edge_index_used"], fx["timestep"], fx["y"], fx["train_mask"],
                               fx["class_weight"], opt, cfg["grad_clip"])
        assert abs(loss - want) <= 1e-6 * abs(want)
    for k, v in model.state_dict().items():
        assert torch.allclose(v.float(), fx["state2"][k].float(), rtol=1e-5, atol=1e-6), k


def test_temporal_masks_reference_fixture():
    """tests/test_masks_and_metrics.py:8-18 of the reference, on the restated mask builder."""
    y = torch.tensor([0, 1, 0, 1, 0])
    t = torch.tensor([1, 1, 2, 3, 4])
    tr, va, te = O.make_temporal_masks(y, t, 1, 3)
    assert tr.tolist() == [True, True, False, False, False]
    assert va.tolist() == [False, False, True, True, False]
    assert te.tolist() == [False, False, False, False, True]
