"""CPU model of the packed 16x2 arithmetic the tile kernel uses in its LAST column pass (gst_kernels.cu, lift_odd_p<LAST>):
the odd outputs of that pass go straight to the (char) truncation `(v & 0x00FF00FF) ^ k`, so the kernel leaves out
  * the bias correction cH = pk(-4096) (a multiple of 256 in each half), and
  * the mask that keeps bit 0 of the high half from being shifted into bit 15 of the low half.
This test checks, over the whole range the wavelet can produce (|x| <= 3488 for any input bytes)
and at its corners, that the two forms agree in the bytes that survive the truncation and that nothing carries from
the low half into the high half."""
import numpy as np

KBIAS = 4096
BOUND = 3488                      # every wavelet intermediate satisfies |x| <= 3488


def pk(v):                        # v in both halves, as the kernel's pk(): (uint32)v * 65537
    return np.uint32(((v & 0xFFFFFFFF) * 65537) & 0xFFFFFFFF)


def pack(lo, hi):
    return (lo.astype(np.uint32) & 0xFFFF) | ((hi.astype(np.uint32) & 0xFFFF) << 16)


def trunc_fix_1(T):               # max(T, min(T + 1, 8193)) per signed 16-bit half
    out = np.zeros_like(T)
    for sh in (0, 16):
        h = ((T >> sh) & 0xFFFF).astype(np.int32)
        h = np.where(h >= 32768, h - 65536, h)
        r = np.maximum(h, np.minimum(h + 1, 8193))
        out |= (r.astype(np.uint32) & 0xFFFF) << sh
    return out


def odd_full(H, EP, EN):          # the general form
    T = (EP + EN) & 0xFFFFFFFF
    T2 = trunc_fix_1(T)
    return ((((T2 & 0xFFFEFFFE) >> 1) + H) + pk(-KBIAS)) & 0xFFFFFFFF


def odd_last(H, EP, EN):          # the LAST form: no correction, no mask
    T = (EP + EN) & 0xFFFFFFFF
    T2 = trunc_fix_1(T)
    return ((T2 >> 1) + H) & 0xFFFFFFFF


def reference_low_bytes(h, ep, en):   # codec/inverse_wavelet.cl:28-64: h + (d0 + d1) / 2 with C division, then (char)
    s = ep + en
    q = np.where(s >= 0, s // 2, -((-s) // 2))
    return (h + q) & 0xFF


def _result(h, ep, en):
    s = ep + en
    return h + np.where(s >= 0, s // 2, -((-s) // 2))


def _check(h_lo, ep_lo, en_lo, h_hi, ep_hi, en_hi):
    # the RESULT is a wavelet intermediate too: only operand triples whose result is within the bound can occur
    # (the general form relies on that as well: its pk(-4096) borrows from the high half otherwise)
    ok = (np.abs(_result(h_lo, ep_lo, en_lo)) <= BOUND) & (np.abs(_result(h_hi, ep_hi, en_hi)) <= BOUND)
    h_lo, ep_lo, en_lo, h_hi, ep_hi, en_hi = (a[ok] for a in (h_lo, ep_lo, en_lo, h_hi, ep_hi, en_hi))
    assert h_lo.size > 1000
    H = pack(h_lo + KBIAS, h_hi + KBIAS)
    EP = pack(ep_lo + KBIAS, ep_hi + KBIAS)
    EN = pack(en_lo + KBIAS, en_hi + KBIAS)
    a, b = odd_full(H, EP, EN), odd_last(H, EP, EN)
    assert np.array_equal(a & 0x00FF00FF, b & 0x00FF00FF)
    # both are the reference's value in the surviving bytes
    assert np.array_equal(b & 0xFF, reference_low_bytes(h_lo, ep_lo, en_lo))
    assert np.array_equal((b >> 16) & 0xFF, reference_low_bytes(h_hi, ep_hi, en_hi))
    # the high half of the LAST form is exact (nothing carried into it): it equals the general form's plus the bias
    assert np.array_equal((b >> 16) & 0xFFFF, (((a >> 16) & 0xFFFF) + KBIAS) & 0xFFFF)


def test_last_column_pass_odd_outputs_random():
    rng = np.random.default_rng(5)
    n = 400000
    v = [rng.integers(-BOUND, BOUND + 1, size=n) for _ in range(6)]
    _check(*v)


def test_last_column_pass_odd_outputs_corners():
    c = np.array([-BOUND, -BOUND + 1, -257, -256, -255, -129, -128, -127, -2, -1, 0, 1, 2, 127, 128, 129, 255, 256, 257,
                  BOUND - 1, BOUND])
    g = np.array(np.meshgrid(c, c, c, c[::3], c[::3], c[::3])).reshape(6, -1)
    _check(*g)
