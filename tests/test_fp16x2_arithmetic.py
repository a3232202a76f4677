"""CPU restatement of the arithmetic the fp16x2 tensor-core kernels use (csrc/s2v_h2.cuh: pow2_scale_for, split2h,
issue_transform_h2 / issue_reduce_h2) -- test infrastructure, not product code.

  x s = h1 + h2 + r,  h1 = RN16(x s),  h2 = RN16(x s - h1),  |r| <= 2^-22 |x s|
  a . b ~= sum_k (h2a h1b + h1a h2b + h1a h1b)   (fp32 accumulation), rescaled by the exact 2^-(ea + eb)

The claims checked here are the ones the GPU tests rely on: the scale puts a tile's largest magnitude in [2^14, 2^15); the split
is exact to 2^-22; a contraction of 64 terms is as accurate as fp32 itself; results do not depend on the power of two chosen as
long as nothing is subnormal; tiles at 1e-30 or 1e+30 neither underflow nor overflow."""
import numpy as np
import pytest
import torch


def pow2_scale_for(amax: float) -> float:
    """csrc/s2v_h2.cuh:pow2_scale_for -- 2^k with amax 2^k in [2^14, 2^15); 2^127 for (near-)zero tiles."""
    e = (np.float32(amax).view(np.int32) >> 23) & 255
    k = int(min(127, max(-126, 141 - int(e))))
    return float(np.ldexp(np.float32(1.0), k))


def split2h(x: torch.Tensor, s: float):
    xs = (x * s).float()
    h1 = xs.half()
    h2 = (xs - h1.float()).half()
    return h1, h2


def dot_h2(a: torch.Tensor, b: torch.Tensor):
    """rows of a [R, K] against columns of b [K, N] with one scale per operand tile; fp32 accumulation (the tensor core's
    accumulator truncates where this rounds: same order of magnitude)."""
    sa, sb = pow2_scale_for(float(a.abs().max())), pow2_scale_for(float(b.abs().max()))
    a1, a2 = split2h(a, sa)
    b1, b2 = split2h(b, sb)
    acc = a2.float() @ b1.float()
    acc = acc + a1.float() @ b2.float()
    acc = acc + a1.float() @ b1.float()
    return acc * (1.0 / sa) * (1.0 / sb)


@pytest.mark.parametrize("amax", [3.3e-30, 1e-8, 0.5, 1.0, 1.9999, 2.0, 150.0, 65504.0, 1e12, 3e38])
def test_scale_puts_the_maximum_in_range(amax):
    s = pow2_scale_for(amax)
    assert np.log2(s) == round(np.log2(s))                      # an exact power of two
    scaled = float(np.float32(amax)) * s
    assert 2.0 ** 14 <= scaled < 2.0 ** 15, (amax, s, scaled)
    assert scaled <= 65504.0                                    # representable in fp16


def test_zero_and_denormal_tiles():
    """Below 2^-112 the scale saturates at 2^127: such tiles use fewer than 22 bits, never more than fp16's range."""
    assert pow2_scale_for(0.0) == 2.0 ** 127
    assert pow2_scale_for(1e-38) == 2.0 ** 127 and float(np.float32(1e-38)) * 2.0 ** 127 < 2.0 ** 15
    assert 0.0 * pow2_scale_for(0.0) == 0.0
    assert np.isfinite(np.float32(1e-45) * np.float32(pow2_scale_for(1e-45)))


@pytest.mark.parametrize("scale", [1e-30, 1e-8, 1.0, 1e4, 1e30])
def test_split_is_exact_to_2_pow_minus_22(scale):
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(64, 64, generator=gen) * scale
    s = pow2_scale_for(float(x.abs().max()))
    h1, h2 = split2h(x, s)
    rec = (h1.double() + h2.double()) / s
    rel = ((rec - x.double()).abs() / x.double().abs().clamp(min=1e-300))
    big = x.abs() >= x.abs().max() * 2.0 ** -18               # entries within 2^18 of the largest keep all 22 bits
    assert float(rel[big].max()) <= 2.0 ** -22
    # the rest is exact to the fp16 subnormal spacing: 2^-24 in scaled units, i.e. <= 2^-38 of the tile's largest entry
    err = (rec - x.double()).abs()
    assert float(err.max()) <= float(x.abs().max()) * 2.0 ** -22
    if bool((~big).any()):
        assert float(err[~big].max()) <= float(x.abs().max()) * 2.0 ** -38


@pytest.mark.parametrize("sa,sb", [(1.0, 1.0), (1e-30, 1.0), (1e-8, 1e3), (1e20, 1e-12)])
def test_contraction_is_fp32_accurate(sa, sb):
    gen = torch.Generator().manual_seed(2)
    a = torch.randn(64, 64, generator=gen) * sa
    b = torch.randn(64, 64, generator=gen) * sb
    ref = a.double() @ b.double()
    err_h2 = float((dot_h2(a, b).double() - ref).abs().max() / ref.abs().max())
    err_f32 = float(((a @ b).double() - ref).abs().max() / ref.abs().max())
    assert err_h2 < 5e-7, err_h2
    assert err_h2 < 4 * err_f32 + 1e-7, (err_h2, err_f32)      # the same class as plain fp32


def test_results_do_not_depend_on_the_power_of_two():
    """Rounding to 11 bits commutes with multiplication by a power of two, so a tile scaled by 2^k or 2^(k-3) gives the same
    bits -- as long as h2 stays a normal fp16 number (entries within 2^18 of the tile's largest)."""
    gen = torch.Generator().manual_seed(3)
    a = torch.rand(64, 64, generator=gen) + 0.5                 # all entries within a factor 3 of each other
    b = torch.rand(64, 64, generator=gen) + 0.5
    outs = []
    for shift in (0, -3, -2):                   # (a larger power would leave the fp16 range: the kernel never does)
        a1, a2 = split2h(a, pow2_scale_for(float(a.max())) * 2.0 ** shift)
        b1, b2 = split2h(b, pow2_scale_for(float(b.max())))
        acc = a2.float() @ b1.float() + a1.float() @ b2.float() + a1.float() @ b1.float()
        outs.append(acc * (1.0 / (pow2_scale_for(float(a.max())) * 2.0 ** shift)))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_rows_far_below_the_tile_maximum_lose_bits_only_in_absolute_terms():
    """A row 2^26 below the tile's largest entry keeps an ABSOLUTE error of <= 2^-22 of the largest entry's products -- the bound
    the parity tests state -- which is why pure row transforms scale per 16 rows and not per 64-row tile."""
    gen = torch.Generator().manual_seed(4)
    a = torch.randn(64, 64, generator=gen)
    a[7] *= 2.0 ** -26
    b = torch.randn(64, 64, generator=gen)
    ref = a.double() @ b.double()
    got = dot_h2(a, b).double()
    assert float((got - ref).abs().max() / ref.abs().max()) < 5e-7
    row_rel = float((got[7] - ref[7]).abs().max() / ref[7].abs().max())
    assert row_rel > 1e-6                                       # relative to ITSELF the small row is worse than fp32 ...
    a16 = a[:16].clone()                                        # ... unless its scale comes from its own 16-row group
    a16[0:7] *= 2.0 ** -26
    a16[8:] *= 2.0 ** -26
    got16 = dot_h2(a16, b).double()
    ref16 = a16.double() @ b.double()
    assert float((got16[7] - ref16[7]).abs().max() / ref16[7].abs().max()) < 5e-7
