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
