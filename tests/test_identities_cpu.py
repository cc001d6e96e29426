"""Integer identities the CUDA kernels rely on, checked in numpy on random and extreme operands (no GPU needed).  Each one replaces a
64-bit or quarter-rate operation of the reference's arithmetic by something cheaper; the GPU parity tests prove the kernels, these
prove the algebra they cite."""
import numpy as np


def _operands(n=200_000, seed=3):
    rng = np.random.default_rng(seed)
    a = rng.integers(-2 ** 31, 2 ** 31, n, dtype=np.int64)
    ext = np.array([-2 ** 31, -2 ** 31 + 1, -2 ** 30, -65537, -65536, -65535, -1, 0, 1, 65535, 65536, 2 ** 30, 2 ** 31 - 2, 2 ** 31 - 1], np.int64)
    return np.concatenate([a, ext])


def _mulhi(a, b):
    """high 32 bits of the signed 64-bit product (IMAD.HI / mul.hi.s32)"""
    return (a.astype(object) * int(b) if np.isscalar(b) else a.astype(object) * b.astype(object)) // (1 << 32)


def test_dc_block_fracmul_is_acc_plus_mulhi():
    """msdr_frontend.cu: FRACMUL_SHL(acc, COEF, 1) = bits [61:30] of acc * COEF (input_adc.cpp:198-212) with COEF = 1048300 << 10:
    4 * COEF = 2^32 - 69 * 2^14, so (acc * COEF) >> 30 = acc + mulhi(acc, -(69 << 14)), and the result fits an int32."""
    coef = 1048300 << 10
    assert 4 * coef == (1 << 32) - (69 << 14)
    acc = _operands()
    want = (acc.astype(object) * coef) >> 30
    got = acc.astype(object) + _mulhi(acc, -(69 << 14))
    assert (want == got).all()
    assert max(want) < 2 ** 31 and min(want) >= -2 ** 31


def test_smlawb_from_16_bit_halves():
    """msdr_device.cuh BqStageS: (c * v) >> 16 for a 32-bit coefficient and an int16 value = ch * v + ((cl * v) >> 16) with
    c = ch * 2^16 + cl, cl unsigned; every partial product fits an int32."""
    rng = np.random.default_rng(5)
    c = _operands(seed=6)
    v = np.concatenate([rng.integers(-32768, 32768, c.size - 4, dtype=np.int64), np.array([-32768, 32767, 0, -1], np.int64)])
    ch, cl = c >> 16, c & 0xFFFF
    want = (c.astype(object) * v.astype(object)) >> 16
    got = (ch * v).astype(object) + ((cl * v) >> 16).astype(object)
    assert (want == got).all()
    assert np.abs(cl * v).max() < 2 ** 31 and np.abs(ch * v).max() <= 2 ** 30


def test_smlawb_is_mulhi_of_the_shifted_value():
    """every integer stage form carries values as v << 16: (c * v) >> 16 = mulhi(c, v << 16) (filter_biquad.cpp:56-63, SMLAWB / SMLAWT)"""
    rng = np.random.default_rng(7)
    c = _operands(seed=8)
    v = rng.integers(-32768, 32768, c.size, dtype=np.int64)
    assert ((c.astype(object) * v.astype(object)) >> 16 == _mulhi(c, v << 16)).all()


def test_ssb_sums_on_packed_words():
    """epilogues: with p = I | Q << 16 the wrapping int16 sums of Minimal-SDR.ino:591-604 are the upper half of
    p * 65537 (USB: I + Q) and of (p ^ 0xFFFF0000) * 65537 + 0x10000 (LSB: I - Q), all mod 2^32."""
    rng = np.random.default_rng(9)
    i = np.concatenate([rng.integers(-32768, 32768, 100_000), [-32768, 32767, -32768, 32767, 0]]).astype(np.int64)
    q = np.concatenate([rng.integers(-32768, 32768, 100_000), [-32768, 32767, 32767, -32768, -32768]]).astype(np.int64)
    p = ((i & 0xFFFF) | ((q & 0xFFFF) << 16)) & 0xFFFFFFFF
    usb = ((p * 65537) & 0xFFFFFFFF) >> 16
    lsb = ((((p ^ 0xFFFF0000) * 65537) + 0x10000) & 0xFFFFFFFF) >> 16
    assert (usb == ((i + q) & 0xFFFF)).all()
    assert (lsb == ((i - q) & 0xFFFF)).all()
    # the same with a byte permute for the shift (msdr_chain_v6.cu): p + (p << 16) instead of p * 65537
    assert ((((p + (p << 16)) & 0xFFFFFFFF) >> 16) == usb).all()


def test_byte_plane_recombination_wraps_like_the_reference_accumulator():
    """tensor-core FIR: x = xh * 256 + xl (xh signed, xl unsigned), same for the taps; the three int32 accumulators hh, hl + lh, ll give
    (hh << 16) + (mid << 8) + ll = sum of x * c mod 2^32, the wrapping 32-bit accumulator of arm_fir_fast_q15.c:60-329."""
    rng = np.random.default_rng(11)
    x = rng.integers(-32768, 32768, (64, 256), dtype=np.int64)
    c = rng.integers(-32768, 32768, (64, 256), dtype=np.int64)
    xh, xl, ch, cl = x >> 8, x & 0xFF, c >> 8, c & 0xFF
    hh, mid, ll = (xh * ch).sum(1), (xh * cl + xl * ch).sum(1), (xl * cl).sum(1)
    for a in (hh, mid, ll):
        assert np.abs(a).max() < 2 ** 31  # each plane accumulator is exact in int32 for 256 taps
    got = ((hh << 16) + (mid << 8) + ll) & 0xFFFFFFFF
    assert (got == ((x * c).sum(1) & 0xFFFFFFFF)).all()
