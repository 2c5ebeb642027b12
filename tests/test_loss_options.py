"""`_make_loss_fn` options (focal, linear / sqrt time weighting, learned-time-table L2; src/train_gnn.py:136-183).
The golden vectors are outputs of the reference's OWN function (tests/golden/make_loss_golden.py).  CPU part: the
oracle restatement vs those vectors.  GPU part: `ops.make_loss_fn` (one kernel: per-row loss, mean, d loss / d logits)
vs the same vectors, and a TrainStep with the options against the oracle step."""
import os

import pytest
import torch

from oracle import pyg_restated as O
from util import REL_FP32, assert_bf16_grads_bounded, assert_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt")


class _Net(torch.nn.Module):
    def __init__(self, emb):
        super().__init__()
        self.time_emb = None
        if emb is not None:
            self.time_emb = torch.nn.Embedding(*emb.shape)
            with torch.no_grad():
                self.time_emb.weight.copy_(emb)


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def test_oracle_loss_matches_reference(gold):
    for c in gold["cases"]:
        net = _Net(c["emb"])
        fn = O.make_loss_fn(c["cfg"], gold["cw"], net, c["t_min"], c["t_max"])
        lg = gold["logits"].clone().requires_grad_(True)
        use_t = c["cfg"].get("time_loss_weighting", "none") != "none"
        loss = fn(lg, gold["target"], gold["t_idx"] if use_t else None)
        loss.backward()
        assert torch.equal(loss.detach(), c["loss"]), c["cfg"]
        assert torch.equal(lg.grad, c["dlogits"]), c["cfg"]
        if c["demb"] is not None:
            assert torch.equal(net.time_emb.weight.grad, c["demb"])


@pytest.mark.gpu
def test_device_loss_matches_reference(gold):
    from egnn_b200 import ops
    for c in gold["cases"]:
        net = _Net(c["emb"]).cuda()
        fn = ops.make_loss_fn(c["cfg"], gold["cw"], net, c["t_min"], c["t_max"])
        lg = gold["logits"].cuda().requires_grad_(True)
        use_t = c["cfg"].get("time_loss_weighting", "none") != "none"
        loss = fn(lg, gold["target"].cuda(), gold["t_idx"].cuda() if use_t else None)
        loss.backward()
        ref = float(c["loss"])
        assert abs(float(loss) - ref) <= REL_FP32 * abs(ref), (c["cfg"], float(loss), ref)
        assert_close(lg.grad, c["dlogits"], REL_FP32, f"dlogits {c['cfg']}")
        if c["demb"] is not None:
            assert_close(net.time_emb.weight.grad, c["demb"], REL_FP32, "d time_emb")
        # the row-indexed form TrainStep uses: the same rows scattered inside a larger logits matrix
        n = lg.size(0)
        big = torch.zeros(3 * n, 2, device="cuda")
        idx = torch.arange(n, device="cuda") * 3 + 1
        big[idx] = lg.detach()
        big.requires_grad_(True)
        yb = torch.full((3 * n,), -1, dtype=torch.int64, device="cuda")
        yb[idx] = gold["target"].cuda()
        tb = torch.zeros(3 * n, dtype=torch.int64, device="cuda")
        tb[idx] = gold["t_idx"].cuda()
        net.zero_grad()
        l2 = fn.on_rows(big, yb, idx, tb if use_t else None)
        l2.backward()
        assert abs(float(l2) - ref) <= REL_FP32 * abs(ref)
        assert_close(big.grad[idx], c["dlogits"], REL_FP32, "dlogits (rows)")
        rest = big.grad.clone()
        rest[idx] = 0
        assert not rest.any()                                                    # zero outside the train rows


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [dict(focal_loss=True, focal_gamma=2.0),
                                  dict(time_loss_weighting="sqrt"),
                                  dict(time_loss_weighting="linear", time_embed_l2=0.01, time_embed_type="learned")])
@pytest.mark.parametrize("amp", [False, True])
def test_train_step_with_loss_options(small_graph, opts, amp):
    """One full step (SAGE-ResBN; fused sequence under bf16, autograd path in fp32) with each option against the oracle
    step driven by the restated `_make_loss_fn`."""
    import egnn_b200 as E
    from egnn_b200 import ops
    from egnn_b200.train import TrainStep
    gr = small_graph
    cfg = dict(hidden_dim=64, layers=3, dropout=0.0, time_embed_dim=2, time_embed_type="sin", max_timestep=49)
    cfg.update(opts)
    ei = torch.cat([gr.edge_index, gr.edge_index.flip(0)], 1)
    torch.manual_seed(1)
    model = E.build_model("sage_resbn", 166, cfg)
    ref = O.build_model("sage_resbn", 166, cfg)
    ref.load_state_dict(model.state_dict())
    model = model.cuda()
    cw = O.class_weight(gr.y[gr.train_mask])
    tt = gr.timestep[gr.train_mask]
    t_min, t_max = int(tt.min()), int(tt.max())
    weighted = opts.get("time_loss_weighting", "none") != "none"
    step = TrainStep(model, gr.x.cuda(), ei.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(), lr=1e-3,
                     weight_decay=5e-5, grad_clip=1.0, amp=amp, loss_fn=ops.make_loss_fn(cfg, cw, model, t_min, t_max))
    loss = float(step.run())
    ours = {n: p.grad.detach().clone().cpu() for n, p in model.named_parameters()}   # views of the flat buffer (pre-clip)
    m = gr.train_mask

    def oracle_grads(net, bf16):
        net.train()
        with torch.autocast(device_type="cpu", dtype=torch.bfloat16, enabled=bf16):
            lg = net(gr.x, ei, gr.timestep)
        fn = O.make_loss_fn(cfg, cw, net, t_min, t_max)
        lt = fn(lg.float()[m], gr.y[m], gr.timestep[m] if weighted else None)
        lt.backward()
        return float(lt), [(n, q.grad) for n, q in net.named_parameters()]

    loss32, g32 = oracle_grads(ref, False)
    if not amp:
        assert abs(loss - loss32) <= REL_FP32 * abs(loss32), (loss, loss32)
        gmax = max(g.abs().max().item() for _, g in g32)
        for n, g in g32:
            if g.abs().max().item() < 1e-5 * gmax:
                continue          # analytically zero (a conv bias feeding BatchNorm): rounding noise on both sides
            e = (ours[n] - g).abs().max().item() / g.abs().max().item()
            assert e <= 2e-5, (n, e)
    else:
        ref16 = O.build_model("sage_resbn", 166, cfg)
        ref16.load_state_dict(ref.state_dict())
        loss16, g16 = oracle_grads(ref16, True)
        assert abs(loss - loss32) <= max(4e-2 * abs(loss32), 3 * abs(loss16 - loss32)), (loss, loss32, loss16)
        assert_bf16_grads_bounded(ours.items(), g32, g16, what=str(opts))
