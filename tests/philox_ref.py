"""Host restatement of the device point generator (Philox4x32-10 + VEGAS grid map) used to cross-check
grb_vegas_iterate against grb_eval.  Layout (graniitti_b200/csrc/grb_point.cuh): key = (seed lo, seed hi),
counter = (point lo, point hi, iteration, draw >> 1); draw d takes words (0,1) if d is even else (2,3);
u = (hi:lo >> 11) * 2^-53.  Draws 0..dim-1 pick the VEGAS bins, the following ones are the internal tape."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32(c, k):
    c0, c1, c2, c3 = c
    k0, k1 = k
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def draw(seed, point, it, d):
    o = philox4x32((point & MASK, (point >> 32) & MASK, it & MASK, d >> 1), (seed & MASK, (seed >> 32) & MASK))
    lo, hi = (o[2], o[3]) if d & 1 else (o[0], o[1])
    return float(((hi << 32) | lo) >> 11) * 2.0 ** -53


def philox_points(xmat, first_call, n, seed, it, ntape):
    bins, dim = xmat.shape
    u = np.zeros((n, dim)); vw = np.ones(n); tape = np.zeros((n, ntape))
    for i in range(n):
        p = first_call + i
        for j in range(dim):
            zn = draw(seed, p, it, j) * bins + 1.0
            ind = max(1, min(int(zn), bins))
            if ind > 1:
                zo = xmat[ind - 1, j] - xmat[ind - 2, j]
                ac = xmat[ind - 2, j] + (zn - ind) * zo
            else:
                zo = xmat[0, j]
                ac = (zn - ind) * zo
            u[i, j] = ac
            vw[i] *= zo * bins
        for t in range(ntape):
            tape[i, t] = draw(seed, p, it, dim + t)
    return u, vw, tape


# ---- vectorised restatement (numpy uint64 lanes) for the full-size parity sets -----------------------------------------
def _philox4x32_np(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on arrays of 32-bit words held in uint64 (so that the 32x32 -> 64 products do not overflow)."""
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ k0) & m, p1 & m, ((p0 >> s32) ^ c3 ^ k1) & m, p0 & m
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def philox_uniforms_np(first_call, n, seed, it, ndraws):
    """[n][ndraws] uniforms of the points first_call .. first_call + n - 1 (same layout as draw())."""
    p = np.arange(first_call, first_call + n, dtype=np.uint64)
    plo, phi = p & np.uint64(MASK), (p >> np.uint64(32)) & np.uint64(MASK)
    itw = np.full(n, it & MASK, dtype=np.uint64)
    out = np.empty((n, ndraws))
    for blk in range((ndraws + 1) // 2):
        o = _philox4x32_np(plo, phi, itw, np.full(n, blk, dtype=np.uint64), seed & MASK, (seed >> 32) & MASK)
        for half in range(2):
            d = 2 * blk + half
            if d >= ndraws:
                break
            lo, hi = (o[2], o[3]) if half else (o[0], o[1])
            out[:, d] = (((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return out


def philox_points_np(xmat, first_call, n, seed, it, ntape):
    """Vectorised philox_points(): (u [n][dim], vegasweight [n], tape [n][ntape]), bit-identical to the scalar loop."""
    bins, dim = xmat.shape
    r = philox_uniforms_np(first_call, n, seed, it, dim + ntape)
    zn = r[:, :dim] * bins + 1.0
    ind = np.clip(zn.astype(np.int64), 1, bins)
    cols = np.arange(dim)[None, :]
    x1 = xmat[ind - 1, cols]
    x0 = np.where(ind > 1, xmat[np.maximum(ind - 2, 0), cols], 0.0)
    zo = np.where(ind > 1, x1 - x0, x1)
    u = np.where(ind > 1, x0 + (zn - ind) * zo, (zn - ind) * zo)
    vw = np.ones(n)
    for j in range(dim):      # same left-to-right product as the device loop
        vw = vw * (zo[:, j] * bins)
    return u, vw, r[:, dim:].copy()
