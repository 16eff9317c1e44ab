"""torch_scatter is an un-vendored dependency of the reference whose semantics the oracle
restates (oracle/scatter_standin.py).  Where the real wheel is importable, cross-check the stand-in
against it once (skipped in this image: torch_scatter is not installed and there is no network)."""
import pytest
import torch

torch_scatter = pytest.importorskip("torch_scatter")


def test_standin_equals_real_torch_scatter():
    from oracle import scatter_standin as S
    gen = torch.Generator().manual_seed(0)
    src = torch.randn(200, 7, generator=gen).relu()            # exact-zero ties like post-ReLU CNN features
    counts = torch.randint(0, 6, (60,), generator=gen)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
    src = src[: int(ptr[-1])]
    for red in ("sum", "mean", "max", "min"):
        a = src.clone().requires_grad_(True)
        b = src.clone().requires_grad_(True)
        ya = torch_scatter.segment_csr(a, ptr, reduce=red)
        yb = S.segment_csr(b, ptr, reduce=red)
        assert torch.allclose(ya, yb, atol=1e-6), red
        w = torch.randn(ya.shape, generator=gen)
        ga, = torch.autograd.grad((ya * w).sum(), a)
        gb, = torch.autograd.grad((yb * w).sum(), b)
        assert torch.allclose(ga, gb, atol=1e-6), red            # first-arg routing for max / min
