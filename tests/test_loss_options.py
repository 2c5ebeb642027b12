"""`_make_loss_fn` options (focal, linear / sqrt time weighting, learned-time-table L2; src/train_gnn.py:136-183).
The golden vectors are outputs of the reference's OWN function (tests/golden/make_loss_golden.py).  CPU part: the
oracle restatement vs those vectors.  GPU part: `ops.make_loss_fn` (one kernel: per-row loss, mean, d loss / d logits)
vs the same vectors, and a TrainStep with the options against the oracle step."""
import os

import pytest
import torch

from oracle import pyg_restated as O
from util import REL_FP32, assert_close

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
    ttype = opts.get("time_embed_type", "sin")
    cfg = dict(hidden_dim=64, layers=3, dropout=0.0, time_embed_dim=2, time_embed_type=ttype, max_timestep=49, **opts)
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
    ref.train()
    with torch.autocast(device_type="cpu", dtype=torch.bfloat16, enabled=amp):
        lg = ref(gr.x, ei, gr.timestep)
    fn = O.make_loss_fn(cfg, cw, ref, t_min, t_max)
    m = gr.train_mask
    loss_t = fn(lg.float()[m], gr.y[m], gr.timestep[m] if weighted else None)
    loss_t.backward()
    loss_ref = float(loss_t)
    tol = 4e-2 if amp else REL_FP32
    assert abs(loss - loss_ref) <= tol * abs(loss_ref), (loss, loss_ref)
    gmax = max(q.grad.abs().max().item() for q in ref.parameters())
    for n, q in ref.named_parameters():
        if q.grad.abs().max().item() < 1e-5 * gmax:
            continue          # analytically zero (a conv bias feeding BatchNorm): rounding noise on both sides
        e = (ours[n] - q.grad).abs().max().item() / q.grad.abs().max().item()
        assert e <= (6e-2 if amp else 2e-5), (n, e)
