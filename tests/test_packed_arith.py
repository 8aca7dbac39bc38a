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


# ---------------------------------------------------------------------------------------------------------------
# The general packed lifting steps (gst_kernels.cu: lift_even_p, lift_odd_p, trunc_fix) against the reference's C
# arithmetic (codec/inverse_wavelet.cl:28-64), for both operand biases the kernel uses: kBias = 4096 for computed
# values, kRaw = 0x1080 for a freshly unpacked coefficient byte.
KRAW = 0x1080


def ph(v):                        # per-half constant, as the kernel's ph()
    return np.uint32(((v & 0xFFFF) * 65537) & 0xFFFFFFFF)


def _halves_s16(w):
    out = []
    for sh in (0, 16):
        h = ((w >> sh) & 0xFFFF).astype(np.int64)
        out.append(np.where(h >= 32768, h - 65536, h))
    return out


def _from_halves(lo, hi):
    return ((lo & 0xFFFF) | ((hi & 0xFFFF) << 16)).astype(np.uint32)


def viaddmin_s16x2(a, b, c):      # min(a + b, c) per signed 16-bit half (wrapping add)
    al, ah = _halves_s16(a)
    bl, bh = _halves_s16(np.broadcast_to(b, a.shape))
    cl, ch = _halves_s16(np.broadcast_to(c, a.shape))
    wrap = lambda x: ((x + 32768) % 65536) - 32768
    return _from_halves(np.minimum(wrap(al + bl), cl), np.minimum(wrap(ah + bh), ch))


def viaddmax_s16x2(a, b, c):
    al, ah = _halves_s16(a)
    bl, bh = _halves_s16(np.broadcast_to(b, a.shape))
    cl, ch = _halves_s16(c)
    wrap = lambda x: ((x + 32768) % 65536) - 32768
    return _from_halves(np.maximum(wrap(al + bl), cl), np.maximum(wrap(ah + bh), ch))


def lift_even_model(S, HP, HN, hb, s_bias):
    cD = pk(2048 + KBIAS - s_bias)
    c, c3 = (ph(2), ph(5)) if hb == KBIAS else (ph(-254), ph(-251))
    U = (HP + HN) & 0xFFFFFFFF
    m = viaddmin_s16x2(U, c3, ph(8192 + 3))
    T3 = viaddmax_s16x2(U, c, m)
    Qc = (((T3 & 0xFFFCFFFC) >> 2) + ((0 - int(cD)) & 0xFFFFFFFF)) & 0xFFFFFFFF
    return (S - Qc) & 0xFFFFFFFF


def c_div(a, k):
    return np.where(a >= 0, a // k, -((-a) // k))


def _triples(rng, n, bound):
    return [rng.integers(-bound, bound + 1, size=n) for _ in range(6)]


def test_even_step_matches_the_reference_for_both_operand_biases():
    rng = np.random.default_rng(11)
    for hb, s_bias, bound_h, bound_s in ((KBIAS, KBIAS, BOUND, BOUND), (KRAW, KBIAS, 128, BOUND), (KRAW, KRAW, 128, 128)):
        s_lo, s_hi = rng.integers(-bound_s, bound_s + 1, size=(2, 300000))
        hp_lo, hn_lo, hp_hi, hn_hi = rng.integers(-bound_h, bound_h + 1, size=(4, 300000))
        want_lo = s_lo - c_div(hp_lo + hn_lo + 2, 4)
        want_hi = s_hi - c_div(hp_hi + hn_hi + 2, 4)
        ok = (np.abs(want_lo) <= BOUND) & (np.abs(want_hi) <= BOUND)
        sel = lambda a: a[ok]
        got = lift_even_model(pack(sel(s_lo) + s_bias, sel(s_hi) + s_bias), pack(sel(hp_lo) + hb, sel(hp_hi) + hb),
                              pack(sel(hn_lo) + hb, sel(hn_hi) + hb), hb, s_bias)
        gl, gh = _halves_s16(got)
        assert np.array_equal(gl - KBIAS, sel(want_lo)) and np.array_equal(gh - KBIAS, sel(want_hi)), (hb, s_bias)


def test_odd_step_matches_the_reference_for_both_operand_biases():
    rng = np.random.default_rng(12)
    for hb, bound_h in ((KBIAS, BOUND), (KRAW, 128)):
        h_lo, h_hi = rng.integers(-bound_h, bound_h + 1, size=(2, 300000))
        ep_lo, en_lo, ep_hi, en_hi = rng.integers(-BOUND, BOUND + 1, size=(4, 300000))
        want_lo = h_lo + c_div(ep_lo + en_lo, 2)
        want_hi = h_hi + c_div(ep_hi + en_hi, 2)
        ok = (np.abs(want_lo) <= BOUND) & (np.abs(want_hi) <= BOUND)
        sel = lambda a: a[ok]
        H = pack(sel(h_lo) + hb, sel(h_hi) + hb)
        EP, EN = pack(sel(ep_lo) + KBIAS, sel(ep_hi) + KBIAS), pack(sel(en_lo) + KBIAS, sel(en_hi) + KBIAS)
        T2 = trunc_fix_1((EP + EN) & 0xFFFFFFFF)
        got = ((((T2 & 0xFFFEFFFE) >> 1) + H) + pk(-hb)) & 0xFFFFFFFF
        gl, gh = _halves_s16(got)
        assert np.array_equal(gl - KBIAS, sel(want_lo)) and np.array_equal(gh - KBIAS, sel(want_hi)), hb
