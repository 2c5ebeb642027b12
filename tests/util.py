import torch

REL_FP32 = 1e-5   # north-star tolerance for fp32 logits and gradients, relative to ||ref||_inf
REL_BF16 = 4e-2   # stated tolerance for the bf16-autocast path (bf16 has an 8-bit mantissa: 2^-8 = 3.9e-3
                  # per rounding; activations pass ~10 roundings through 3 layers + BatchNorm)


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    denom = ref.abs().max().item()
    if denom == 0:
        return (got - ref).abs().max().item()
    return (got - ref).abs().max().item() / denom


def assert_close(got, ref, rel, what=""):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    e = rel_err(got, ref)
    assert e <= rel, f"{what}: rel err {e:.3e} > {rel:.1e}"
    return e


def assert_close_gated(got, ref, rel, what="", max_units=2, loose=1e-3):
    """`assert_close` for the gradient of a parameter that sits below a ReLU, at sizes where a gate is bound to sit on
    its threshold: of the 13 M layer-0 pre-activations of rec_k8 on the full graph, ~10 lie within 1e-6 of zero in the
    fp32 oracle (profiles/r02/f32_tc_probe.txt), i.e. within the distance between two CORRECT fp32 evaluations of the
    layer (different summation orders; 3xTF32 products are ~1e-6 apart from FFMA ones).  Such a gate evaluated the
    other way changes dz in ONE (row, channel) and therefore exactly ONE output channel of that layer's weight / bias
    gradients (a rank-1 difference, shown by the probe) -- it is a discontinuity of the function, not an error of the
    kernel.  So: every output channel (row of a 2-D gradient, element of a 1-D one) is held to `rel` of ||ref||_inf,
    except that at most `max_units` of them may miss it, and those are held to `loose`.  Returns the worst error among
    the channels held to `rel` and the number of excused channels."""
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    g, r = got.detach().double().cpu(), ref.detach().double().cpu()
    denom = max(r.abs().max().item(), 1e-300)
    err = (g - r).abs()
    per_unit = (err.flatten(1).max(dim=1).values if err.dim() > 1 else err.flatten()) / denom
    bad = per_unit > rel
    n_bad = int(bad.sum())
    assert n_bad <= max_units, f"{what}: {n_bad} output channels miss rel {rel:.1e} (worst {per_unit.max():.3e})"
    if n_bad:
        assert per_unit[bad].max().item() <= loose, f"{what}: excused channel at {per_unit[bad].max():.3e} > {loose:.1e}"
    ok = per_unit[~bad]
    return (ok.max().item() if ok.numel() else 0.0), n_bad


def assert_bf16_grads_bounded(named_ours, named_ref32, named_ref16, what="", factor=3.0, floor=1.5 * REL_BF16):
    """bf16-autocast gradients, bounded RELATIVE TO THE bf16 ORACLE'S OWN DISTANCE FROM fp32: for every
    parameter, ||g_ours - g_fp32||_inf <= max(factor * ||g_oracle_bf16 - g_fp32||_inf, floor * ||g_fp32||_inf).
    Our kernels accumulate in fp32 and round once where PyG's bf16 path accumulates in bf16, so `factor` only has to
    cover different rounding points, not a looser algorithm; the floor (6 % of ||g_fp32||_inf: a gradient passes
    through the forward AND the backward roundings, 1.5 x the forward tolerance) covers tiny tensors (a 2-element
    attention vector) on which the bf16 oracle's own error happens to be small.  Tensors whose fp32 gradient is analytically zero (a conv
    bias feeding BatchNorm) are compared against the global gradient scale.  Returns the worst ratio seen."""
    ref32 = {n: g.detach().double().cpu() for n, g in named_ref32}
    ref16 = {n: g.detach().double().cpu() for n, g in named_ref16}
    gmax = max(g.abs().max().item() for g in ref32.values())
    worst = 0.0
    for n, g in named_ours:
        g = g.detach().double().cpu()
        assert torch.isfinite(g).all(), f"{what} grad {n} not finite"
        scale = max(ref32[n].abs().max().item(), 1e-5 * gmax)
        e_ours = (g - ref32[n]).abs().max().item()
        e_or = (ref16[n] - ref32[n]).abs().max().item()
        bound = max(factor * e_or, floor * scale)
        assert e_ours <= bound, (f"{what} grad {n}: |ours-fp32| {e_ours:.3e} > bound {bound:.3e} "
                                 f"(bf16 oracle |.-fp32| {e_or:.3e}, scale {scale:.3e})")
        worst = max(worst, e_ours / bound)
    return worst
