"""Layer- and net-level parity of the drop-in GCNConv / SAGEConv / GATConv and the four nets
against the restated PyG oracle on identical seeded inputs and copied state dicts:
fp32 logits and gradients within rel 1e-5 (of ||ref||_inf), bf16 autocast within REL_BF16."""
import copy

import pytest
import torch

from oracle import pyg_restated as O
from util import REL_BF16, REL_FP32, assert_bf16_grads_bounded, assert_close, rel_err

pytestmark = pytest.mark.gpu


def _pair(make_ours, make_ref):
    torch.manual_seed(0)
    ours = make_ours()
    ref = make_ref()
    ref.load_state_dict(ours.state_dict())
    return ours.cuda(), ref


def _grads_close(ours, ref, tol, tag):
    worst = 0.0
    for (n1, p1), (n2, p2) in zip(ours.named_parameters(), ref.named_parameters()):
        assert n1 == n2
        assert p1.grad is not None, n1
        worst = max(worst, assert_close(p1.grad, p2.grad, tol, f"{tag} grad {n1}"))
    return worst


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("dims", [(166, 64), (167, 128), (64, 2), (12, 8)])
def test_sage_conv(egnn, small_graph, sym, dims):
    fi, fo = dims
    gr = small_graph
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1) if sym else gr.edge_index
    ours, ref = _pair(lambda: egnn.SAGEConv(fi, fo), lambda: O.SAGEConv(fi, fo))
    torch.manual_seed(1)
    x = torch.randn(gr.num_nodes, fi)
    xr = x.clone().requires_grad_(True)
    xc = x.cuda().requires_grad_(True)
    yr = ref(xr, ei)
    yo = ours(xc, ei.cuda())
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g.cuda())
    assert_close(yo, yr, REL_FP32, "sage out")
    assert_close(xc.grad, xr.grad, REL_FP32, "sage dx")
    _grads_close(ours, ref, REL_FP32, "sage")


@pytest.mark.parametrize("dims", [(167, 128), (128, 2), (10, 5)])
def test_gcn_conv(egnn, small_graph, dims):
    fi, fo = dims
    gr = small_graph
    ei = gr.edge_index
    ours, ref = _pair(lambda: egnn.GCNConv(fi, fo), lambda: O.GCNConv(fi, fo))
    with torch.no_grad():
        ours.bias.uniform_(-1, 1)
        ref.bias.copy_(ours.bias.cpu())
    torch.manual_seed(1)
    x = torch.randn(gr.num_nodes, fi)
    xr, xc = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    yr, yo = ref(xr, ei), ours(xc, ei.cuda())
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g.cuda())
    assert_close(yo, yr, REL_FP32, "gcn out")
    assert_close(xc.grad, xr.grad, REL_FP32, "gcn dx")
    _grads_close(ours, ref, REL_FP32, "gcn")


@pytest.mark.parametrize("cfg", [(167, 8, 4, True), (32, 2, 1, False), (20, 3, 2, False), (16, 5, 3, True)])
def test_gat_conv(egnn, small_graph, cfg):
    fi, c, h, concat = cfg
    gr = small_graph
    ei = gr.edge_index
    ours, ref = _pair(lambda: egnn.GATConv(fi, c, heads=h, concat=concat),
                      lambda: O.GATConv(fi, c, heads=h, concat=concat))
    with torch.no_grad():
        ours.bias.uniform_(-1, 1)
        ref.bias.copy_(ours.bias.cpu())
    torch.manual_seed(1)
    x = torch.randn(gr.num_nodes, fi)
    xr, xc = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    yr, yo = ref(xr, ei), ours(xc, ei.cuda())
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g.cuda())
    assert_close(yo, yr, REL_FP32, "gat out")
    assert_close(xc.grad, xr.grad, 2 * REL_FP32, "gat dx")
    _grads_close(ours, ref, 2 * REL_FP32, "gat")


def test_gat_zero_attention_is_mean_with_self(egnn):
    """all-zero att_* => alpha uniform = 1/(in-deg+1)."""
    ei = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]]).cuda()
    conv = egnn.GATConv(5, 5, heads=1).cuda()
    with torch.no_grad():
        conv.lin.weight.copy_(torch.eye(5))
        conv.att_src.zero_()
        conv.att_dst.zero_()
    out = conv(torch.eye(5).cuda(), ei).cpu()
    want = torch.eye(5)
    for i in range(1, 5):
        want[i] = 0.5 * (torch.eye(5)[i] + torch.eye(5)[i - 1])
    assert torch.allclose(out, want, atol=1e-7)


def test_sage_known_answers(egnn):
    """SAGE mean on the 5-node path graph with x = I."""
    from egnn_b200 import ops, _lib
    ei = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]]).cuda()
    x = torch.nn.functional.pad(torch.eye(5), (0, 3)).cuda()
    g = egnn.build_graph(ei, 5)
    m = ops.spmm(g, "csr", _lib.SPMM_MEAN, x, torch.float32).cpu()[:, :5]
    want = torch.zeros(5, 5)
    for i in range(1, 5):
        want[i, i - 1] = 1
    assert torch.equal(m, want)
    gs = egnn.build_graph(ei, 5, symmetrize=True)
    m = ops.spmm(gs, "csr", _lib.SPMM_MEAN, x, torch.float32).cpu()[:, :5]
    want = torch.zeros(5, 5)
    want[0, 1] = 1
    want[4, 3] = 1
    for i in (1, 2, 3):
        want[i, i - 1] = want[i, i + 1] = 0.5
    assert torch.equal(m, want)


CONFIGS = {
    "gcn": dict(arch="gcn", in_dim=167, hidden_dim=128, layers=3, dropout=0.5, sym=False, lr=3e-3, wd=1e-4),
    "sage": dict(arch="sage", in_dim=167, hidden_dim=128, layers=2, dropout=0.5, sym=True, lr=3e-3, wd=1e-4),
    "rec_k8": dict(arch="sage_resbn", in_dim=166, hidden_dim=64, layers=3, dropout=0.2, sym=True, lr=5e-4,
                   wd=5e-5, time_embed_dim=2, time_embed_type="sin", max_timestep=49),
    "gat": dict(arch="gat", in_dim=167, hidden_dim=32, layers=2, heads=4, dropout=0.5, sym=False, lr=3e-3,
                wd=1e-4),
    "sage_l3": dict(arch="sage", in_dim=167, hidden_dim=128, layers=3, dropout=0.4, sym=True, lr=1e-3, wd=5e-5),
}


def _inputs(gr, cfg):
    x = gr.x
    if cfg["in_dim"] == 167:  # use_time_scalar: train_gnn.py:315-317
        x = torch.cat([x, (gr.timestep.float() / float(gr.timestep.max())).unsqueeze(1)], dim=1)
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1) if cfg["sym"] else gr.edge_index
    return x, ei


@pytest.mark.parametrize("name", list(CONFIGS))
def test_net_eval_logits(egnn, small_graph, name):
    cfg = CONFIGS[name]
    gr = small_graph
    x, ei = _inputs(gr, cfg)
    ours, ref = _pair(lambda: egnn.build_model(cfg["arch"], cfg["in_dim"], cfg),
                      lambda: O.build_model(cfg["arch"], cfg["in_dim"], cfg))
    ours.eval()
    ref.eval()
    with torch.no_grad():
        lo = ours(x.cuda(), ei.cuda(), gr.timestep.cuda())
        lr_ = ref(x, ei, gr.timestep)
    assert_close(lo, lr_, REL_FP32, f"{name} eval logits")


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("with_dropout", [False, True])
def test_net_train_step_fp32(egnn, small_graph, name, with_dropout):
    """One full train_epoch body (forward all nodes, masked weighted CE, backward, clip 1.0, Adam):
    loss, logits, every parameter gradient, BN running stats and the updated parameters.  With
    dropout > 0 the oracle is fed the very keep-masks the CUDA path drew."""
    from egnn_b200 import ops
    from egnn_b200.train import TrainStep
    cfg = dict(CONFIGS[name])
    if not with_dropout:
        cfg["dropout"] = 0.0
    gr = small_graph
    x, ei = _inputs(gr, cfg)
    ours, ref = _pair(lambda: egnn.build_model(cfg["arch"], cfg["in_dim"], cfg),
                      lambda: O.build_model(cfg["arch"], cfg["in_dim"], cfg))
    ours.set_dropout_seed(1234)
    cw = O.class_weight(gr.y[gr.train_mask])
    step = TrainStep(ours, x.cuda(), ei.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(),
                     lr=cfg["lr"], weight_decay=cfg["wd"], grad_clip=1.0, amp=False, cw=cw)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=cfg["lr"], weight_decay=cfg["wd"])
    for it in range(2):
        if it > 0:  # re-synchronise so every iteration checks the kernels, not the Adam-amplified drift
            ref.load_state_dict({k: v.detach().cpu() for k, v in ours.state_dict().items()})
        loss_o = step.run()
        grads_o = {n: p.grad.detach().clone() for n, p in ours.named_parameters()}
        masks = None
        if with_dropout:
            hid = cfg["hidden_dim"]
            masks = [ops.dropout_mask(gr.num_nodes, hid, cfg["dropout"], 1234, li,
                                      seed_off=ours._drop.offset).cpu() for li in range(cfg["layers"] - 1)]
        # reference step, keeping the pre-clip gradients for comparison
        ref.train()
        opt_ref.zero_grad(set_to_none=True)
        uses_t = getattr(ref, "time_embed_dim", 0) > 0
        kw = {} if masks is None else {"dropout_masks": masks}
        logits_r = ref(x, ei, gr.timestep if uses_t else None, **kw)
        loss_r = O.masked_weighted_ce(logits_r, gr.y, gr.train_mask, cw)
        loss_r.backward()
        assert_close(loss_o, loss_r, REL_FP32, f"{name} loss it{it}")
        gmax = max(p.grad.abs().max().item() for p in ref.parameters())
        for n, p in ref.named_parameters():
            if p.grad.abs().max().item() < 1e-5 * gmax:
                # analytically-zero gradient (a conv bias feeding BatchNorm): both sides hold only
                # rounding noise of a cancelling sum; compare against the global gradient scale
                assert grads_o[n].abs().max().item() < 1e-4 * gmax, f"{name} grad {n} it{it}"
                continue
            assert_close(grads_o[n], p.grad, 2 * REL_FP32, f"{name} grad {n} it{it}")
        gref = {n: p.grad.detach().clone() for n, p in ref.named_parameters()}
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        opt_ref.step()
        # Adam's update g/(|g|+eps) is scale-free per element, so an element whose gradient is tiny
        # next to ||g||_inf (and therefore carries a large RELATIVE error at rel 1e-5 of ||g||_inf)
        # moves by a visibly different amount; compare where |g| >= 1% of ||g||_inf (the Adam kernel
        # itself is checked against torch on identical gradients in test_clip_adam_matches_torch)
        for (n, p), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
            d = (p.data.cpu() - pr.data).abs()
            assert d.max().item() <= 2.0 * cfg["lr"] * (it + 1), f"{name} param {n} it{it}"
            big = gref[n].abs() >= 1e-2 * max(gref[n].abs().max().item(), 1e-5 * gmax)
            if big.any() and it == 0:
                assert d[big].max().item() <= 2e-3 * cfg["lr"], f"{name} param {n} it{it}: {d[big].max().item():.3e}"
    if cfg["arch"] == "sage_resbn":
        for b, br in zip(ours.bns, ref.bns):
            assert_close(b.running_mean, br.running_mean, REL_FP32, "running_mean")
            assert_close(b.running_var, br.running_var, REL_FP32, "running_var")
            assert int(b.num_batches_tracked) == int(br.num_batches_tracked) == 2


@pytest.mark.parametrize("name", ["rec_k8", "sage", "gcn", "gat"])
def test_net_train_step_bf16_autocast(egnn, small_graph, name):
    """bf16 autocast step against the oracle under CPU autocast(bf16).  Stated tolerance REL_BF16:
    our kernels accumulate in fp32 and round once where PyG accumulates in bf16 (strictly tighter)."""
    from egnn_b200.train import TrainStep
    cfg = dict(CONFIGS[name])
    cfg["dropout"] = 0.0
    gr = small_graph
    x, ei = _inputs(gr, cfg)
    ours, ref = _pair(lambda: egnn.build_model(cfg["arch"], cfg["in_dim"], cfg),
                      lambda: O.build_model(cfg["arch"], cfg["in_dim"], cfg))
    ref32 = copy.deepcopy(ref)
    cw = O.class_weight(gr.y[gr.train_mask])
    step = TrainStep(ours, x.cuda(), ei.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(),
                     lr=cfg["lr"], weight_decay=cfg["wd"], grad_clip=1.0, amp=True, cw=cw)
    loss_o = step.run()
    grads_o = [(n, p.grad.detach().clone()) for n, p in ours.named_parameters()]      # pre-clip gradients
    opt = torch.optim.Adam(ref.parameters(), lr=cfg["lr"], weight_decay=cfg["wd"])
    loss_r, _ = O.train_step(ref, x, ei, gr.timestep, gr.y, gr.train_mask, cw, opt, 0.0,
                             amp_dtype=torch.bfloat16)
    opt32 = torch.optim.Adam(ref32.parameters(), lr=cfg["lr"], weight_decay=cfg["wd"])
    loss_32, _ = O.train_step(ref32, x, ei, gr.timestep, gr.y, gr.train_mask, cw, opt32, 0.0)
    assert abs(float(loss_o) - loss_r) <= REL_BF16 * abs(loss_r)
    # we must be at least as close to the fp32 truth as the bf16 oracle is (plus slack)
    assert abs(float(loss_o) - loss_32) <= abs(loss_r - loss_32) + REL_BF16 * abs(loss_32)
    # gradients: bounded by the bf16 oracle's own distance from fp32 (grad_clip 0 above keeps the oracles' raw grads)
    assert_bf16_grads_bounded(grads_o, [(n, p.grad) for n, p in ref32.named_parameters()],
                              [(n, p.grad) for n, p in ref.named_parameters()], name)


def test_rec_k8_cuda_graph_replay_matches_eager(egnn, small_graph):
    from egnn_b200.train import TrainStep
    cfg = CONFIGS["rec_k8"]
    gr = small_graph
    x, ei = _inputs(gr, cfg)
    outs = []
    for graphed in (False, True):
        torch.manual_seed(0)
        m = egnn.build_model(cfg["arch"], cfg["in_dim"], cfg).cuda()
        m.set_dropout_seed(77)
        st = TrainStep(m, x.cuda(), ei.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(),
                       lr=cfg["lr"], weight_decay=cfg["wd"], amp=False)
        if graphed:
            st.capture(warmup=2)     # 2 eager warm-up steps + 1 captured-but-not-run
            losses = [float(st.run()) for _ in range(3)]
        else:
            losses = [float(st.run()) for _ in range(5)][2:]
        outs.append((losses, [p.detach().clone() for p in m.parameters()]))
    (l0, p0), (l1, p1) = outs
    assert l0 == l1, (l0, l1)           # deterministic kernels: bitwise equal trajectories
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)


def _degenerate_graphs():
    from egnn_b200 import synthetic
    path = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 4]])          # the reference's own fixture (tests/test_masks_and_metrics.py:12)
    none = torch.zeros((2, 0), dtype=torch.int64)               # no edges at all: every node isolated
    adv = synthetic.adversarial_tiny().edge_index               # duplicates, reciprocal pair, double self-loop, hub of 500
    return {"path5": (5, path), "no_edges": (7, none), "adversarial": (509, adv)}


@pytest.mark.parametrize("gname", ["path5", "no_edges", "adversarial"])
@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("amp", [False, True])
def test_nets_on_degenerate_graphs(egnn, name, gname, amp):
    """Edge cases through every net and both precisions: the 5-node path graph, a graph without edges (mean over an
    empty neighbourhood = 0, GCN/GAT reduce to the self-loop), and the adversarial multigraph.  Logits in train mode
    with dropout 0 (BatchNorm uses batch statistics) and all parameter gradients against the oracle."""
    cfg = dict(CONFIGS[name], dropout=0.0)
    n, ei = _degenerate_graphs()[gname]
    torch.manual_seed(3)
    x = torch.randn(n, cfg["in_dim"])
    t = torch.randint(1, 50, (n,))
    ours, ref = _pair(lambda: egnn.build_model(cfg["arch"], cfg["in_dim"], cfg),
                      lambda: O.build_model(cfg["arch"], cfg["in_dim"], cfg))
    ours.train()
    ref.train()
    g = torch.randn(n, 2)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        lo = ours(x.cuda(), ei.cuda(), t.cuda())
    lr_ = ref(x, ei, t)
    (lo.float() * g.cuda()).sum().backward()
    (lr_ * g).sum().backward()
    ref16 = None
    if amp:
        ref16 = copy.deepcopy(ref)
        ref16.zero_grad(set_to_none=True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            l16 = ref16(x, ei, t)
        (l16.float() * g).sum().backward()
    tol = REL_BF16 if amp else 2 * REL_FP32
    assert_close(lo, lr_, tol, f"{name}/{gname} logits")
    gmax = max(p.grad.abs().max().item() for p in ref.parameters())
    if amp:
        for _, p1 in ours.named_parameters():
            assert torch.isfinite(p1.grad).all()
        if n >= 100:   # 5-7 nodes: one ReLU gate flipped by bf16 rounding moves a gradient by O(1/n); logits are the check
            assert_bf16_grads_bounded([(k, p.grad) for k, p in ours.named_parameters()],
                                      [(k, p.grad) for k, p in ref.named_parameters()],
                                      [(k, p.grad) for k, p in ref16.named_parameters()], f"{name}/{gname}")
        return
    for (n1, p1), (_, p2) in zip(ours.named_parameters(), ref.named_parameters()):
        assert torch.isfinite(p1.grad).all(), n1
        if p2.grad.abs().max().item() < 1e-5 * gmax:
            continue  # analytically zero (conv bias in front of BatchNorm): rounding noise on both sides
        assert_close(p1.grad, p2.grad, 2.5 * tol, f"{name}/{gname} grad {n1}")


class _MiniRefStyle(torch.nn.Module):
    """A net written the way the reference writes them (`src/models/gnn.py`): library convs with ordinary torch ops
    -- nn.BatchNorm1d, F.relu, F.dropout, a residual nn.Linear -- in between.  Only the conv class differs."""

    def __init__(self, convs_mod, kind):
        super().__init__()
        mk = {"sage": lambda i, o: convs_mod.SAGEConv(i, o), "gcn": lambda i, o: convs_mod.GCNConv(i, o),
              "gat": lambda i, o: convs_mod.GATConv(i, o // 4, heads=4) if o > 2 else convs_mod.GATConv(i, o, heads=1, concat=False)}[kind]
        self.c1, self.c2, self.c3 = mk(166, 64), mk(64, 64), mk(64, 2)
        self.bn = torch.nn.BatchNorm1d(64)
        self.res = torch.nn.Linear(166, 64, bias=False)

    def forward(self, x, ei):
        h = torch.nn.functional.relu(self.bn(self.c1(x, ei))) + self.res(x)
        h = torch.nn.functional.dropout(torch.nn.functional.elu(self.c2(h, ei)), p=0.0, training=self.training)
        return self.c3(h, ei)


@pytest.mark.parametrize("kind", ["sage", "gcn", "gat"])
@pytest.mark.parametrize("amp", [False, True])
def test_convs_drop_into_a_reference_style_module(egnn, small_graph, kind, amp):
    """INTEGRATION.md's one-line switch: `from egnn_b200 import GCNConv, SAGEConv, GATConv` inside a module that
    otherwise uses plain torch ops (autograd, nn.BatchNorm1d, autocast) -- logits and gradients against the same
    module built on the oracle convs."""
    gr = small_graph
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    torch.manual_seed(5)
    ours = _MiniRefStyle(egnn, kind)
    ref = _MiniRefStyle(O, kind)
    ref.load_state_dict(ours.state_dict())
    ours = ours.cuda().train()
    ref.train()
    x = gr.x
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        lo = ours(x.cuda(), ei.cuda())
    lr_ = ref(x, ei)
    w = torch.randn(lr_.shape, generator=torch.Generator().manual_seed(9))
    (lo.float() * w.cuda()).sum().backward()
    (lr_ * w).sum().backward()
    tol = REL_BF16 if amp else 2 * REL_FP32
    assert_close(lo, lr_, tol, f"{kind} logits")
    if amp:
        ref16 = copy.deepcopy(ref)
        ref16.zero_grad(set_to_none=True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            l16 = ref16(x, ei)
        (l16.float() * w).sum().backward()
        assert_bf16_grads_bounded([(k, p.grad) for k, p in ours.named_parameters()],
                                  [(k, p.grad) for k, p in ref.named_parameters()],
                                  [(k, p.grad) for k, p in ref16.named_parameters()], kind)
        return
    for (n1, p1), (_, p2) in zip(ours.named_parameters(), ref.named_parameters()):
        b = p2.grad.detach().double().flatten()
        if b.abs().max() < 1e-5 * max(q.grad.abs().max() for q in ref.parameters()):
            continue
        assert_close(p1.grad, p2.grad, 5 * REL_FP32, f"{kind} grad {n1}")


def test_backward_after_a_no_grad_forward_accumulates_exactly(egnn, small_graph):
    """fp32 tensor-core GEMMs accumulate in TMEM for no-grad forwards and exactly for everything a gradient is taken of
    (`egnn_set_f32_tc_exact`): an `eval_split`-style forward between a graph's forward and its backward must not leave
    the backward's dgrad / wgrad products in the truncating mode."""
    from egnn_b200 import _lib
    L = _lib.lib()
    gr = small_graph
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    ours, ref = _pair(lambda: egnn.SAGEConv(168, 64), lambda: O.SAGEConv(168, 64))
    torch.manual_seed(4)
    x = torch.randn(gr.num_nodes, 168)
    xr, xc = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    yo = ours(xc, ei.cuda())
    with torch.no_grad():
        ours(xc, ei.cuda())
    assert L.egnn_set_f32_tc_exact(0) == 0          # the no-grad forward switched to the in-TMEM accumulate
    g = torch.randn(gr.num_nodes, 64)
    yo.backward(g.cuda())
    assert L.egnn_set_f32_tc_exact(1) == 1          # ... and the backward switched back before its first GEMM
    ref(xr, ei).backward(g)
    assert_close(xc.grad, xr.grad, REL_FP32, "sage dx")
    _grads_close(ours, ref, REL_FP32, "sage")
