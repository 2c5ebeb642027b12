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


RANK_KEYS = ("pr_auc", "n", "positives", "thresholds", "roc_auc", "_5", "_6", "_7", "max_f1", "threshold_max_f1",
             "precision_at_k", "recall_at_precision", "threshold_for_precision", "f1_at_threshold", "ece", "k_used")


def ranking_metrics(y: torch.Tensor, mask: Optional[torch.Tensor] = None, *, logits: Optional[torch.Tensor] = None,
                    scores: Optional[torch.Tensor] = None, top_k: int = 100, target_precision: float = 0.90,
                    threshold: Optional[torch.Tensor] = None, ece_bins: int = 15,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Everything `src/train_gnn.py:449-470` computes from (y, scores) over the rows selected by `mask`, on the
    device, from ONE sort: float64[16] laid out as `RANK_KEYS` (PR-AUC, ROC-AUC, max-F1 and its threshold,
    precision@k, recall@precision, the threshold reaching `target_precision`, F1 at `threshold`, ECE).
    `threshold`: a float64 DEVICE tensor (first element is used), e.g. `val_out[9:10]` -- the reference picks the
    threshold on the validation rows and scores the test rows with it (`:455-466`); None = this selection's own
    max-F1 threshold (`use_val_for_thresholds: false`)."""
    if (logits is None) == (scores is None):
        raise ValueError("give exactly one of logits / scores")
    src = logits if logits is not None else scores
    if not src.is_cuda:
        raise RuntimeError("egnn_b200 computes metrics on the GPU only (no CPU fallback)")
    if src.dtype != torch.float32 or y.dtype != torch.int64:
        raise TypeError("logits / scores must be float32 and y int64")
    n = int(y.numel())
    if logits is not None and (logits.dim() != 2 or logits.size(1) < 2 or logits.size(0) != n or logits.stride(1) != 1):
        raise ValueError("logits must be [N, >=2] with contiguous rows")
    if scores is not None and (scores.numel() != n or not scores.is_contiguous()):
        raise ValueError("scores must be a contiguous [N] tensor")
    if mask is not None and (mask.numel() != n or mask.dtype not in (torch.bool, torch.uint8)
                             or not mask.is_contiguous()):
        raise ValueError("mask must be a contiguous bool / uint8 [N] tensor")
    if threshold is not None and (threshold.dtype != torch.float64 or not threshold.is_cuda):
        raise TypeError("threshold must be a float64 CUDA tensor")
    if out is None:
        out = torch.empty(16, dtype=torch.float64, device=src.device)
    elif out.dtype != torch.float64 or out.numel() < 16 or not out.is_contiguous() or not out.is_cuda:
        raise ValueError("out must be a contiguous float64 CUDA tensor with at least 16 elements")
    L = lib()
    ws_bytes = L.egnn_ap_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=src.device)
    check(L.egnn_ranking_metrics(ptr(logits), logits.stride(0) if logits is not None else 0, ptr(scores),
                                 ptr(y.contiguous()), ptr(mask), n, int(top_k), float(target_precision),
                                 ptr(threshold), int(ece_bins), None, ptr(out), ptr(ws), ws_bytes, stream()))
    return out


def fit_temperature(logits: torch.Tensor, y: torch.Tensor, mask: Optional[torch.Tensor] = None,
                    max_iter: int = 100) -> torch.Tensor:
    """`TemperatureScaler().fit(logits_val[val_mask], y[val_mask])` (`src/utils/calibrate.py:8-30`,
    `src/train_gnn.py:424-429`) on the device -> float64[5] = [T, NLL at T=1, NLL at T, iterations, rows]; divide the
    logits by `out[0]` before the softmax as `get_probs` does (`:431-442`)."""
    if not logits.is_cuda:
        raise RuntimeError("egnn_b200 fits the temperature on the GPU only (no CPU fallback)")
    if logits.dtype != torch.float32 or logits.dim() != 2 or logits.size(1) != 2 or logits.stride(1) != 1:
        raise TypeError("logits must be float32 [N, 2] with contiguous rows")
    if y.dtype != torch.int64 or y.numel() != logits.size(0):
        raise TypeError("y must be int64 [N]")
    if mask is not None and (mask.numel() != y.numel() or mask.dtype not in (torch.bool, torch.uint8)
                             or not mask.is_contiguous()):
        raise ValueError("mask must be a contiguous bool / uint8 [N] tensor")
    out = torch.empty(5, dtype=torch.float64, device=logits.device)
    check(lib().egnn_temperature_fit(ptr(logits), logits.stride(0), ptr(y.contiguous()), ptr(mask), y.numel(),
                                     int(max_iter), ptr(out), stream()))
    return out


@torch.no_grad()
def final_metrics(model, x, edge_index, timestep, y, val_mask, test_mask, *, calibrate_temperature: bool = True,
                  top_k: int = 100, precision_target: Optional[float] = None,
                  use_val_for_thresholds: bool = True) -> dict:
    """The reference's run tail (`src/train_gnn.py:420-480`) with nothing leaving the device until the final read:
    optional temperature scaling on the validation rows, one fp32 eval forward, the decision threshold picked on the
    validation rows (max F1, or the lowest threshold reaching `precision_target`), PR-AUC / ROC-AUC / F1 at that
    threshold / precision@k / recall@precision / ECE on the test rows.  `precision_target=None` means the key is absent
    from the config (`cfg.get("precision_target", 0.0)` for the threshold picker, `cfg.get("precision_target", 0.90)`
    for recall@precision, `:457,471`).  Returns the reference's `metrics` dict keys."""
    from .train import model_uses_time_embed
    model.eval()
    logits = model(x, edge_index, timestep if model_uses_time_embed(model) else None).float().contiguous()
    T = None
    if calibrate_temperature:
        T = fit_temperature(logits, y, val_mask)
        logits = (logits / T[0].float()).contiguous()       # `logits / ts.T` (fp32 parameter in the reference)
    rp_target = 0.90 if precision_target is None else float(precision_target)
    val = ranking_metrics(y, val_mask, logits=logits, top_k=top_k, target_precision=rp_target)
    if use_val_for_thresholds:
        thr = val[12:13] if (precision_target and precision_target > 0) else val[9:10]
    else:
        thr = None
    test = ranking_metrics(y, test_mask, logits=logits, top_k=top_k, target_precision=rp_target, threshold=thr)
    v, t = val.tolist(), test.tolist()                        # the one host synchronisation
    thr_used = (v[12] if (precision_target and precision_target > 0) else v[9]) if use_val_for_thresholds else t[9]
    return {"pr_auc_illicit": t[0], "roc_auc": t[4], "f1_illicit_at_thr": t[13], "threshold": thr_used,
            "precision_at_k": t[10], "recall_at_precision": t[11], "ece": t[14], "n_test": int(t[1]),
            "val_pr_auc": v[0], "val_max_f1": v[8], "temperature": float(T[0]) if T is not None else None}


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
                                       self.patience, stream()))
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
        capture: bool = True, log=None, health_check=None, loss_fn=None) -> dict:
    """The reference's full-batch training loop (`/root/reference/src/train_gnn.py:375-417`) with the epoch tail on
    the device: per epoch one (captured) train step, one fp32 eval forward, the validation PR-AUC and the
    early-stopping update -- no host synchronisation.  The host looks at the device state every `poll_every`
    epochs (`if bad >= patience: break`, `:411`).  The best parameters are snapshotted on the device at the epoch they
    occur and the device state freezes once `bad >= patience` (`egnn_early_stop_update`), so the restored model is
    the reference's `best_state` even though the loop runs up to `poll_every - 1` steps past the stopping epoch.
    Epoch 1 is the first optimizer step from the initial weights: the CUDA-graph warm-up steps are rolled back
    (`TrainStep.capture(preserve_state=True)`).  `health_check` (e.g. `ShardedContext.check`) runs at every poll.
    `loss_fn`: an `ops.make_loss_fn(cfg, cw, model, t_min, t_max)` callable for the focal / time-weighted / embed-L2
    variants of `_make_loss_fn` (`src/train_gnn.py:136-183,372`); default = the class-weighted CE.
    Returns {best_val, best_epoch, epochs, stop_epoch, loss}."""
    from .train import TrainStep
    step = TrainStep(model, x, edge_index, timestep, y, train_mask, lr=lr, weight_decay=weight_decay,
                     grad_clip=grad_clip, amp=amp, health_check=health_check, loss_fn=loss_fn)
    ev = None
    if capture:
        from .train import EvalStep
        step.capture(warmup=2, preserve_state=True)
        ev = EvalStep(model, x, edge_index, timestep).capture()     # the per-epoch eval forward as one graph too
    buffers = [b for b in model.buffers() if b.is_cuda and b.numel() > 0]
    stopper = EarlyStopper(patience=patience, flat_param=step.opt.flat_param, extra=buffers)
    ap = torch.empty(8, dtype=torch.float64, device=x.device)
    epochs = 0
    for epoch in range(1, max_epochs + 1):
        loss = step.run()
        if ev is not None:
            average_precision(y, val_mask, logits=ev.run()[1].float().contiguous(), out=ap)
        else:
            eval_pr_auc(model, x, edge_index, timestep, y, val_mask, out=ap)
        stopper.update(ap)
        epochs = epoch
        if epoch % poll_every == 0 or epoch == max_epochs:
            st = stopper.state.tolist()                     # the one host synchronisation per `poll_every` epochs
            if health_check is not None:
                health_check()
            if log is not None:
                log(epoch, float(loss), float(ap[0]), st[0])
            if st[1] >= patience:
                break
    st = stopper.state.tolist()
    stopper.restore_best()
    # st[3] = epochs the device bookkeeping saw before it froze = the epoch the reference's loop breaks at
    return {"best_val": st[0], "best_epoch": int(st[2]), "epochs": epochs, "stop_epoch": int(st[3]),
            "loss": float(step.loss)}
