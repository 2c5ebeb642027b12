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
