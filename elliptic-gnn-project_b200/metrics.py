"""Epoch tail on the device: validation PR-AUC and early stopping without host round trips.

Mirrors what the reference does on the host after every epoch (`/root/reference/src/train_gnn.py:387-411`):
`eval_split` -> `pr_auc_illicit((y_val == 1).astype(int), p_val)` (`/root/reference/src/utils/metrics.py:11-13`,
sklearn `average_precision_score`) -> `if pr_val > best_val: best_state = clone(state_dict)` / `bad += 1` ->
`if bad >= patience: break`.  Here the score, the comparison and the best-parameter snapshot stay on the GPU
(`csrc/metrics.cu`); the host reads `EarlyStopper.state` only when it decides to (e.g. every 10 epochs).
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import check, lib, ptr, stream


def average_precision(y: torch.Tensor, mask: Optional[torch.Tensor] = None, *, logits: Optional[torch.Tensor] = None,
                      scores: Optional[torch.Tensor] = None, scores_out: Optional[torch.Tensor] = None,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PR-AUC of the class `y == 1` over the rows selected by `mask` -> float64[8] ON THE DEVICE:
    `[AP, selected rows, positives, distinct thresholds, ROC-AUC, 0, 0, 0]`.  Give `logits` ([N, 2] fp32; the score is
    `softmax(logits, 1)[:, 1]` as in `eval_split`) or `scores` ([N] fp32)."""
    if (logits is None) == (scores is None):
        raise ValueError("give exactly one of logits / scores")
    src = logits if logits is not None else scores
    if not src.is_cuda:
        raise RuntimeError("egnn_b200 computes metrics on the GPU only (no CPU fallback)")
    if src.dtype != torch.float32:
        raise TypeError("logits / scores must be float32 (eval forwards are never under autocast)")
    if y.dtype != torch.int64:
        raise TypeError("y must be int64")
    n = int(y.numel())
    if logits is not None:
        if logits.dim() != 2 or logits.size(1) < 2 or logits.size(0) != n or logits.stride(1) != 1:
            raise ValueError("logits must be [N, >=2] with contiguous rows")
    elif scores.numel() != n or not scores.is_contiguous():
        raise ValueError("scores must be a contiguous [N] tensor")
    if mask is not None:
        if mask.numel() != n or mask.dtype not in (torch.bool, torch.uint8) or not mask.is_contiguous():
            raise ValueError("mask must be a contiguous bool / uint8 [N] tensor")
    if out is None:
        out = torch.empty(8, dtype=torch.float64, device=src.device)
    elif out.dtype != torch.float64 or out.numel() < 8 or not out.is_contiguous() or not out.is_cuda:
        raise ValueError("out must be a contiguous float64 CUDA tensor with at least 8 elements")
    L = lib()
    ws_bytes = L.egnn_ap_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=src.device)
    check(L.egnn_average_precision(ptr(logits), logits.stride(0) if logits is not None else 0, ptr(scores),
                                   ptr(y.contiguous()), ptr(mask), n, ptr(scores_out), ptr(out), ptr(ws), ws_bytes,
                                   stream()))
    return out


class EarlyStopper:
    """`best_val` / `bad` / `best_state` of `/root/reference/src/train_gnn.py:375-411` as device state.

    `state` (float64[5], device) = [best value, epochs since the best, epoch of the best, epochs seen, improved].
    `update(ap)` enqueues the comparison and, when `flat_param` was given, the conditional snapshot of the flat
    parameter buffer (`train.FlatClipAdam.flat_param`) into `best_param`; nothing is read back.  `should_stop()`
    is the one host synchronisation (`bad >= patience`)."""

    def __init__(self, patience: int = 20, flat_param: Optional[torch.Tensor] = None, device=None, extra=()):
        """`extra`: further tensors of the state dict to snapshot with the parameters (BatchNorm running_mean /
        running_var / num_batches_tracked live outside the flat parameter buffer)."""
        dev = flat_param.device if flat_param is not None else torch.device(device or "cuda")
        self.patience = int(patience)
        self.state = torch.tensor([-1.0, 0.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
        self.flat_param = flat_param
        self.best_param = torch.empty_like(flat_param) if flat_param is not None else None
        self.extra = [t for t in extra if t.numel() > 0]
        for t in self.extra:
            if not t.is_contiguous() or (t.numel() * t.element_size()) % 4:
                raise ValueError("extra tensors must be contiguous with a 4-byte multiple size")
        self.best_extra = [torch.empty_like(t) for t in self.extra]

    def update(self, ap: torch.Tensor) -> None:
        n = int(self.flat_param.numel()) if self.flat_param is not None else 0
        L = lib()
        check(L.egnn_early_stop_update(ptr(ap), ptr(self.state), ptr(self.flat_param), ptr(self.best_param), n,
                                       stream()))
        for t, b in zip(self.extra, self.best_extra):
            check(L.egnn_snapshot_if_improved(ptr(self.state), ptr(t), ptr(b), t.numel() * t.element_size(), stream()))

    def should_stop(self) -> bool:
        return float(self.state[1].item()) >= self.patience

    @property
    def best(self) -> float:
        return float(self.state[0].item())

    def restore_best(self) -> None:
        """`model.load_state_dict(best_state)` (`src/train_gnn.py:416-417`) for the flat parameter buffer."""
        if float(self.state[2].item()) > 0:
            if self.flat_param is not None:
                self.flat_param.copy_(self.best_param)
            for t, b in zip(self.extra, self.best_extra):
                t.copy_(b)


@torch.no_grad()
def eval_pr_auc(model, x, edge_index, timestep, y, mask, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`eval_split` + `pr_auc_illicit` (`src/train_gnn.py:248-257,391`): fp32 eval forward, then the PR-AUC of the
    illicit class over `mask`, all on the device."""
    from .train import model_uses_time_embed
    model.eval()
    logits = model(x, edge_index, timestep if model_uses_time_embed(model) else None)
    return average_precision(y, mask, logits=logits.float().contiguous(), out=out)


def fit(model, x, edge_index, timestep, y, train_mask, val_mask, *, lr: float, weight_decay: float,
        grad_clip: float = 1.0, amp: bool = False, max_epochs: int = 200, patience: int = 20, poll_every: int = 10,
        capture: bool = True, log=None) -> dict:
    """The reference's full-batch training loop (`/root/reference/src/train_gnn.py:375-417`) with the epoch tail on
    the device: per epoch one (captured) train step, one fp32 eval forward, the validation PR-AUC and the
    early-stopping update -- no host synchronisation.  The host looks at the device state every `poll_every`
    epochs (`if bad >= patience: break`, `:411`); because the best parameters are snapshotted on the device at the
    epoch they occur, the restored model is the reference's `best_state` even when the loop overshoots the stopping
    epoch by up to `poll_every - 1` steps.  Returns {best_val, best_epoch, epochs, loss}."""
    from .train import TrainStep
    step = TrainStep(model, x, edge_index, timestep, y, train_mask, lr=lr, weight_decay=weight_decay,
                     grad_clip=grad_clip, amp=amp)
    step.run()
    if capture:
        step.capture(warmup=1)
    buffers = [b for b in model.buffers() if b.is_cuda and b.numel() > 0]
    stopper = EarlyStopper(patience=patience, flat_param=step.opt.flat_param, extra=buffers)
    ap = torch.empty(8, dtype=torch.float64, device=x.device)
    epochs = 0
    for epoch in range(1, max_epochs + 1):
        loss = step.run()
        eval_pr_auc(model, x, edge_index, timestep, y, val_mask, out=ap)
        stopper.update(ap)
        epochs = epoch
        if epoch % poll_every == 0 or epoch == max_epochs:
            st = stopper.state.tolist()                     # the one host synchronisation per `poll_every` epochs
            if log is not None:
                log(epoch, float(loss), float(ap[0]), st[0])
            if st[1] >= patience:
                break
    st = stopper.state.tolist()
    stopper.restore_best()
    return {"best_val": st[0], "best_epoch": int(st[2]), "epochs": epochs, "loss": float(step.loss)}
