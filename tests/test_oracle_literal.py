"""CPU: the loop-for-loop restatement (what `bench.py --impl reference` times) is bit-equal to
the vectorised oracle on small inputs."""
import torch

from oracle import fusion_ref as fr
from oracle import literal as lit


def test_literal_pair_and_lloyd_bitequal(books):
    g = torch.Generator().manual_seed(41)
    d3 = torch.exp(0.3 * torch.randn(1, 1, 8, 8, generator=g))
    assert torch.equal(lit.pair_v1_literal(d3), fr.pair_v1(d3))
    x = torch.exp(0.3 * torch.randn(1, 1, 16, 16, generator=g))
    dn_1 = fr.resize_half(x)
    raw = lit.pair_id_literal(x, dn_1)
    assert raw.dtype == torch.float64 and torch.equal(raw, fr.pair_id(x, dn_1))
    sub = raw[:, :8].contiguous()
    assert torch.equal(lit.lloyd_literal(sub, *books[16]), fr.lloyd(sub, *books[16])[0])
    sub32 = fr.pair_v1(d3)[:, :8].contiguous()
    assert torch.equal(lit.lloyd_literal(sub32, *books[8]), fr.lloyd(sub32, *books[8])[0])


def test_literal_tail_8(books):
    g = torch.Generator().manual_seed(42)
    x = torch.exp(0.3 * torch.randn(2, 1, 8, 8, generator=g))
    assert torch.equal(lit.relative_decoder_tail_literal(x, books), fr.relative_decoder_tail(x, books))
