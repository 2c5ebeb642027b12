"""fp32 training on the tensor cores: per-parameter gradient error of ONE full-size train step against the CPU oracle.
Usage: [EGNN_F32_TC_TRAIN=1] [EGNN_F32_TC_WGRAD=1] [EGNN_TF32_DEBUG=4] python profiles/f32_tc_probe.py [config ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

import egnn_b200 as egnn
from egnn_b200 import ops, synthetic
from egnn_b200.train import TrainStep
from oracle import pyg_restated as O
from test_gpu_convs import CONFIGS, _inputs, _pair


def main():
    names = sys.argv[1:] or ["rec_k8"]
    gr = synthetic.make_elliptic_like(train_window_k=8)
    for name in names:
        cfg = dict(CONFIGS[name])
        x, ei = _inputs(gr, cfg)
        ours, ref = _pair(lambda: egnn.build_model(cfg["arch"], cfg["in_dim"], cfg),
                          lambda: O.build_model(cfg["arch"], cfg["in_dim"], cfg))
        ours.set_dropout_seed(2024)
        cw = O.class_weight(gr.y[gr.train_mask])
        step = TrainStep(ours, x.cuda(), ei.cuda(), gr.timestep.cuda(), gr.y.cuda(), gr.train_mask.cuda(),
                         lr=cfg["lr"], weight_decay=cfg["wd"], grad_clip=1.0, amp=False, cw=cw)
        loss_o = float(step.run())
        grads_o = {n: p.grad.detach().clone().cpu() for n, p in ours.named_parameters()}
        masks = [ops.dropout_mask(gr.num_nodes, cfg["hidden_dim"], cfg["dropout"], 2024, li,
                                  seed_off=ours._drop.offset).cpu() for li in range(cfg["layers"] - 1)]
        ref.train()
        uses_t = getattr(ref, "time_embed_dim", 0) > 0
        loss_r = O.masked_weighted_ce(ref(x, ei, gr.timestep if uses_t else None, dropout_masks=masks), gr.y,
                                      gr.train_mask, cw)
        loss_r.backward()
        print(f"== {name}: TC_TRAIN={ops.F32_TC_TRAIN} TC_WGRAD={ops.F32_TC_WGRAD} "
              f"TF32_DEBUG={os.environ.get('EGNN_TF32_DEBUG', '0')} fused={step._fused_ok()} "
              f"loss rel {abs(loss_o - float(loss_r)) / abs(float(loss_r)):.2e}")
        for n, p in ref.named_parameters():
            r = p.grad
            e = (grads_o[n] - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
            print(f"   {n:28s} |ref|max {r.abs().max().item():.3e}  rel err {e:.3e}")
        if os.environ.get("PROBE_GATES") and hasattr(ref, "bns"):
            # is the layer-0 error ONE flipped ReLU gate?  (a pre-activation within rounding distance of 0 whose gate
            # differs between two fp32 evaluations changes dz0 in a single (row, column): a rank-1 gradient difference)
            for key in ("convs.0.lin_r.weight", "convs.0.lin_l.weight"):
                D = grads_o[key] - dict(ref.named_parameters())[key].grad
                rn = D.norm(dim=1)
                top = torch.topk(rn, 4)
                print(f"   {key}: row norms of the difference, top 4: "
                      + ", ".join(f"col {int(i)}: {float(v):.3e}" for v, i in zip(top.values, top.indices))
                      + f"; median {float(rn.median()):.3e}")
            pre = {}
            hk = ref.bns[0].register_forward_hook(lambda m, i, o: pre.__setitem__("a", o.detach()))
            with torch.no_grad():
                ref(x, ei, gr.timestep if uses_t else None, dropout_masks=masks)
            hk.remove()
            a = pre["a"]
            for tau in (1e-7, 1e-6, 1e-5, 1e-4):
                print(f"   oracle BN-0 outputs with |a| < {tau:g}: {int((a.abs() < tau).sum())} of {a.numel()}")
        # step time (eager launches are host-bound: CUDA graph)
        step.capture(warmup=2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            step.run()
        e1.record()
        torch.cuda.synchronize()
        print(f"   step {e0.elapsed_time(e1) / 20:.4f} ms (CUDA graph)")


if __name__ == "__main__":
    main()
