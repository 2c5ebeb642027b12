"""Golden vectors for the loss options, produced by the REFERENCE's own `_make_loss_fn`
(/root/reference/src/train_gnn.py:136-183; torch_geometric stubbed exactly as in make_golden.py, the loss never touches
it): loss value and d loss / d logits for plain / focal / linear / sqrt / embed-L2 combinations on seeded logits.
    python tests/golden/make_loss_golden.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (its import-time side effects: sys.path + helpers only)

MG._stub()
from src import train_gnn as R  # noqa: E402


class Net(torch.nn.Module):
    def __init__(self, learned):
        super().__init__()
        self.time_emb = torch.nn.Embedding(49, 4) if learned else None


cases = []
g = torch.Generator().manual_seed(7)
n = 700
logits = (torch.randn(n, 2, generator=g) * 3.0)
logits[:5] = torch.tensor([[40.0, -40.0], [-40.0, 40.0], [0.0, 0.0], [1e-3, -1e-3], [25.0, 24.0]])
target = (torch.rand(n, generator=g) < 0.2).long()
t_idx = torch.randint(20, 35, (n,), generator=g)
cw = R.class_weight(target)
for cfg in [dict(), dict(focal_loss=True), dict(focal_loss=True, focal_gamma=1.0), dict(focal_loss=True, focal_gamma=0.5),
            dict(time_loss_weighting="linear"), dict(time_loss_weighting="sqrt"),
            dict(focal_loss=True, focal_gamma=3.0, time_loss_weighting="sqrt"),
            dict(time_loss_weighting="linear", time_embed_l2=0.05, learned=True),
            dict(time_embed_l2=0.3, learned=True)]:
    cfg = dict(cfg)
    learned = cfg.pop("learned", False)
    for t_min, t_max in [(27, 34), (34, 34)]:
        torch.manual_seed(3)
        net = Net(learned)
        fn = R._make_loss_fn(cfg, cw, net, t_min, t_max)
        lg = logits.clone().requires_grad_(True)
        use_t = cfg.get("time_loss_weighting", "none") != "none"
        loss = fn(lg, target, t_idx if use_t else None)
        loss.backward()
        cases.append({"cfg": cfg, "learned": learned, "t_min": t_min, "t_max": t_max, "loss": loss.detach().clone(),
                      "dlogits": lg.grad.clone(),
                      "emb": net.time_emb.weight.detach().clone() if learned else None,
                      "demb": net.time_emb.weight.grad.clone() if learned and net.time_emb.weight.grad is not None else None})
out = os.path.join(HERE, "loss_golden.pt")
torch.save({"logits": logits, "target": target, "t_idx": t_idx, "cw": cw, "cases": cases}, out)
print("wrote", out, len(cases), "cases", os.path.getsize(out), "bytes")
