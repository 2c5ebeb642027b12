"""fp16 autocast (what the reference's `amp: true` selects on CUDA: `torch.autocast` default dtype + GradScaler,
`/root/reference/src/train_gnn.py:36-47,202-207`).  There are no fp16 kernels: an fp16-autocast region is computed in
fp32 (`ops.amp_bf16`, `ops.widen_fp16`), so the results are the fp32 path's -- inside the reference's own fp16 rounding
error -- and GradScaler sees ordinary fp32 gradients.  `EGNN_FP16_AUTOCAST=raise` keeps the old strict behaviour."""
import warnings

import pytest
import torch

from util import REL_FP32, assert_close


def _fake_autocast(monkeypatch, enabled, dtype):
    monkeypatch.setattr(torch, "is_autocast_enabled", lambda *a, **k: enabled)
    monkeypatch.setattr(torch, "get_autocast_dtype", lambda *a, **k: dtype)


def test_autocast_policy_host_logic(monkeypatch, egnn):
    """The dtype switch every conv / net entry goes through, without a GPU (autocast state faked)."""
    from egnn_b200 import ops
    monkeypatch.delenv("EGNN_FP16_AUTOCAST", raising=False)
    _fake_autocast(monkeypatch, False, torch.float16)
    assert ops.amp_bf16() is False
    _fake_autocast(monkeypatch, True, torch.bfloat16)
    assert ops.amp_bf16() is True
    _fake_autocast(monkeypatch, True, torch.float16)
    monkeypatch.setattr(ops, "_FP16_WARNED", False)
    with pytest.warns(RuntimeWarning, match="computed in fp32"):
        assert ops.amp_bf16() is False                       # widened, and said so once
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert ops.amp_bf16() is False                       # ... only once
    h = torch.arange(6, dtype=torch.float16).view(2, 3)
    w = ops.widen_fp16(h)
    assert w.dtype == torch.float32 and torch.equal(w, h.float())
    x = torch.zeros(2, 3)
    assert ops.widen_fp16(x) is x and ops.widen_fp16(x.bfloat16()).dtype == torch.bfloat16
    monkeypatch.setenv("EGNN_FP16_AUTOCAST", "raise")
    with pytest.raises(RuntimeError, match="bf16 autocast only"):
        ops.amp_bf16()
    with pytest.raises(TypeError, match="float16"):
        ops.widen_fp16(h)
    monkeypatch.setenv("EGNN_FP16_AUTOCAST", "fp64")
    with pytest.raises(ValueError):
        ops.amp_bf16()
    _fake_autocast(monkeypatch, True, torch.float64)
    monkeypatch.delenv("EGNN_FP16_AUTOCAST")
    with pytest.raises(RuntimeError, match="bf16 autocast only"):
        ops.amp_bf16()


NETS = {
    "gcn": dict(arch="gcn", in_dim=166, hidden_dim=64, layers=3, dropout=0.0),
    "gat": dict(arch="gat", in_dim=166, hidden_dim=32, layers=2, dropout=0.0, heads=4),
    "rec_k8": dict(arch="sage_resbn", in_dim=166, hidden_dim=64, layers=3, dropout=0.0, use_bn=True, residual=True,
                   time_embed_dim=2, time_embed_type="sin", max_timestep=49),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(NETS))
def test_fp16_autocast_region_is_the_fp32_path(egnn, small_graph, name, monkeypatch):
    """One train-mode forward / backward (dropout 0) under `torch.autocast('cuda')` (fp16, the reference's amp context)
    against the same call without autocast: fp32 logits, same values, same gradients."""
    monkeypatch.delenv("EGNN_FP16_AUTOCAST", raising=False)
    cfg, gr = NETS[name], small_graph
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
    x, t = gr.x.cuda(), gr.timestep.cuda()
    torch.manual_seed(11)
    net = egnn.build_model(cfg["arch"], cfg["in_dim"], cfg).cuda().train()
    w = torch.randn(x.size(0), 2, device="cuda", generator=torch.Generator("cuda").manual_seed(2))

    def run(amp):
        net.zero_grad(set_to_none=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                lg = net(x, ei, t)
        (lg.float() * w).sum().backward()
        return lg.detach(), [(n, p.grad.detach().clone()) for n, p in net.named_parameters() if p.grad is not None]

    if name == "rec_k8":      # BatchNorm running statistics advance per call: same starting point for both runs
        state = {k: v.clone() for k, v in net.state_dict().items()}
    l32, g32 = run(False)
    if name == "rec_k8":
        net.load_state_dict(state)
    l16, g16 = run(True)
    assert l16.dtype == torch.float32
    assert_close(l16, l32, REL_FP32, f"{name} logits under fp16 autocast")
    assert len(g16) == len(g32) > 0
    for (n, a), (_, b) in zip(g16, g32):
        assert a.dtype == torch.float32
        if b.abs().max() > 0:
            assert_close(a, b, REL_FP32, f"{name} grad {n} under fp16 autocast")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sage", "gcn", "gat"])
def test_fp16_activations_and_gradscaler(egnn, small_graph, kind, monkeypatch):
    """The reference's training idiom (`scaler.scale(loss).backward(); scaler.unscale_; clip; scaler.step`) around a
    module whose own `nn.Linear` hands the conv an fp16 tensor: the conv computes on the widened values, returns the
    upstream module an fp16 gradient, and the scaler finds no inf / nan (its scale stays put)."""
    monkeypatch.delenv("EGNN_FP16_AUTOCAST", raising=False)
    gr = small_graph
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1).cuda()
    x = gr.x.cuda()
    torch.manual_seed(13)
    pre = torch.nn.Linear(166, 64).cuda()
    conv = {"sage": lambda: egnn.SAGEConv(64, 32), "gcn": lambda: egnn.GCNConv(64, 32),
            "gat": lambda: egnn.GATConv(64, 8, heads=4)}[kind]().cuda()
    opt = torch.optim.Adam(list(pre.parameters()) + list(conv.parameters()), lr=1e-3)
    scaler = torch.amp.GradScaler("cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        with torch.autocast("cuda", dtype=torch.float16):
            h = pre(x)
            assert h.dtype == torch.float16
            h.retain_grad()
            out = conv(h, ei)
    assert out.dtype == torch.float32
    ref = conv(h.detach().float(), ei)                        # the same conv on the widened activations, no autocast
    assert_close(out, ref, REL_FP32, f"{kind} conv on fp16 activations")
    loss = out.float().pow(2).mean()
    scale0 = scaler.get_scale()
    scaler.scale(loss).backward()
    assert h.grad is not None and h.grad.dtype == torch.float16
    scaler.unscale_(opt)
    torch.nn.utils.clip_grad_norm_(list(pre.parameters()) + list(conv.parameters()), 1.0)
    for p in list(pre.parameters()) + list(conv.parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all()
    before = [p.detach().clone() for p in conv.parameters()]
    scaler.step(opt)
    scaler.update()
    assert scaler.get_scale() == scale0                       # no overflow was found, the step was taken
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, conv.parameters()))


@pytest.mark.gpu
def test_fp16_autocast_strict_mode_raises(egnn, small_graph, monkeypatch):
    monkeypatch.setenv("EGNN_FP16_AUTOCAST", "raise")
    gr = small_graph
    conv = egnn.SAGEConv(166, 32).cuda()
    ei = gr.edge_index.cuda()
    with torch.autocast("cuda", dtype=torch.float16):
        with pytest.raises(RuntimeError, match="bf16 autocast only"):
            conv(gr.x.cuda(), ei)
    with pytest.raises(TypeError, match="float16"):
        conv(gr.x.cuda().half(), ei)
