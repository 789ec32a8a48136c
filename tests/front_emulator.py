"""Lane-level NumPy model of the warp-synchronous banded LDL^T front that
csrc/vbfem_front.cuh implements (test infrastructure, CPU only).

It mirrors the kernel's index logic one to one -- 32 lanes with cyclic row
ownership, a sliding register window whose slot t always holds column j+t of
the lane's row (every rank-1 update writes its result one slot down, so the
window shifts for free and the column loop needs no unrolling), the rank-1
update applied one column late (look-ahead), four static reload slots for the
lanes between two rows, the mirrored bottom front, the Schur dump / merge at
the shared middle block, and the extra right-hand sides that give the
observed-node displacement as a dot product -- so that the slot / lane
arithmetic and the algebra of the merged forward+adjoint solve are validated
on a CPU against dense solves, independently of the CUDA code.
"""
import numpy as np

W = 32
RELOAD_SLOTS = (5, 12, 19, 26)   # static slots; entry distance e = k - t, each e in 0..25 exactly once


class Front:
    """One elimination front in LOCAL coordinates: ``band[c, t] = A[c+t, c]`` for the
    ``ncols`` columns stored for this front; rows run to ``nrows`` (> ncols for the
    bottom front, whose trailing rows are the shared middle block)."""

    def __init__(self, band, ncols, nrows, b, z):
        self.band, self.ncols, self.nrows, self.b = band, ncols, nrows, b
        self.P = b + 1
        self.NS = self.P + 1                     # 27 slots
        self.z = z                               # list of rhs vectors (len >= ncols), updated in place
        nr = len(z)
        lane = np.arange(W)
        self.k = lane.copy()
        self.R = lane.copy()
        self.cur = np.zeros((W, self.NS))
        self.zr = np.zeros((W, nr))
        for ln in range(W):
            for t in range(self.NS):
                self.cur[ln, t] = self._entry(ln, 0, t)
            for r in range(nr):
                self.zr[ln, r] = z[r][ln] if ln < ncols else 0.0
        self.vp = np.zeros(W)
        self.wp = np.zeros(self.P)
        self.ydot = np.zeros(nr)
        self.bad = False

    def _entry(self, ln, j, t):
        """Pristine band entry of lane ln's row at column j+t (0 when it does not exist)."""
        R, c = self.R[ln], j + t
        e = R - c
        if 0 <= e <= self.b and R < self.nrows and c < self.ncols:
            return self.band[c, e]
        return 0.0

    def step(self, j):
        b, P = self.b, self.P
        k, R = self.k, self.R
        act = (k <= b) & (R < self.nrows)
        v = np.where(act, self.cur[:, 0], 0.0)
        src = j & 31
        d, v1, zj = v[src], v[(src + 1) & 31], self.zr[src].copy()
        # pending rank-1 update of column j-1, written one slot down (the window shift)
        for t in range(1, b):
            self.cur[:, t - 1] = self.cur[:, t] - self.vp * self.wp[t + 1]
        self.cur[:, b - 1] = self.cur[:, b]
        self.cur[:, b] = self.cur[:, b + 1]
        if not d > 0.0:
            self.bad = True
        rd = 1.0 / d
        self.cur[:, 0] -= (v * v1) * rd
        w = v * rd
        for ln in range(W):
            if k[ln] <= b:
                self.band[j, k[ln]] = rd if k[ln] == 0 else w[ln]
        for r in range(len(self.z)):
            self.z[r][j] = zj[r]
        self.wp = self.band[j].copy()
        # the pivot lane starts reloading its next row at the end of this step: its own (dead)
        # row must not receive the late update
        self.vp = np.where(k >= 1, v, 0.0)
        wz = np.where((k >= 1) & (k <= b), w, 0.0)
        self.zr -= wz[:, None] * zj[None, :]
        self.ydot += zj[0] * rd * zj
        # advance
        wrap = k == 0
        self.k = (k - 1) & 31
        self.R = R + 32 * wrap
        k, R = self.k, self.R
        for ln in range(W):
            if not (k[ln] >= b and R[ln] < self.nrows):
                continue
            for t in RELOAD_SLOTS:
                e = k[ln] - t
                if 0 <= e <= b:
                    assert R[ln] - e == j + 1 + t
                    # columns >= ncols (bottom front: the middle-middle block, owned by the top
                    # front) read as zero: the kernel keeps a zeroed region behind the band
                    self.cur[ln, t] = self.band[j + 1 + t, e] if j + 1 + t < self.ncols else 0.0
        for ln in range(W):
            if wrap[ln]:
                for r in range(len(self.z)):
                    self.zr[ln, r] = self.z[r][R[ln]] if R[ln] < self.ncols else 0.0

    def flush(self):
        for t in range(1, self.b):
            self.cur[:, t] -= self.vp * self.wp[t + 1]
        self.vp[:] = 0.0

    def dump(self, j):
        """Bottom front after its ncols columns: Schur contributions to the P middle rows."""
        P = self.P
        S, zs = np.zeros((P, P)), np.zeros((P, self.zr.shape[1]))
        for ln in range(W):
            kk = self.k[ln]
            if kk < P:
                assert self.R[ln] == j + kk
                for t in range(kk + 1):
                    S[kk, t] = self.cur[ln, t]
                zs[kk] = self.zr[ln]
        return S, zs

    def merge(self, j, S, zs):
        """Top front at its first middle column: add the mirrored contributions."""
        P = self.P
        for ln in range(W):
            kk = self.k[ln]
            if kk < P:
                assert self.R[ln] == j + kk
                for t in range(kk + 1):
                    self.cur[ln, t] += S[P - 1 - t, P - 1 - kk]
                self.zr[ln] += zs[P - 1 - kk]


def twisted_forward_adjoint(A, f, b, pT, tip, gy, w_mid_fn):
    """Factor A (SPD, half bandwidth b) with the twisted scheme and return
    (u, y, psi): u = A^-1 f, y = u[tip] (two rows of the bottom front, obtained as
    dot products of forward-eliminated vectors), psi = A^-1 (gy[0] e_tip0 + gy[1]
    e_tip1 + w_mid) where w_mid = w_mid_fn(u_mid) is supported on the middle rows."""
    n = A.shape[0]
    P = b + 1
    nB = n - pT - P
    assert nB >= P and pT >= P
    me = pT + P
    bandT = np.zeros((me, P))
    bandB = np.zeros((nB, P))
    for hi in range(n):
        for lo in range(max(0, hi - b), hi + 1):
            if hi < me:
                bandT[lo, hi - lo] = A[hi, lo]
            else:
                bandB[n - 1 - hi, hi - lo] = A[hi, lo]
    # local vectors: [top | middle | bottom (mirrored)]
    zT = [np.array(f[:me], dtype=float), np.zeros(me), np.zeros(me)]
    zB = [np.array(f[::-1][:nB], dtype=float), np.zeros(nB), np.zeros(nB)]
    for i in range(2):
        zB[1 + i][n - 1 - tip[i]] = 1.0
    T = Front(bandT, me, me, b, zT[:1])
    Bf = Front(bandB, nB, nB + P, b, zB)
    for j in range(pT):
        T.step(j)
    for j in range(nB):
        Bf.step(j)
    Bf.flush()
    S, zs = Bf.dump(nB)
    T.flush()
    # the top front continues with three right-hand sides
    T.z = zT
    T.zr = np.concatenate([T.zr, np.zeros((W, 2))], axis=1)
    T.ydot = np.concatenate([T.ydot, np.zeros(2)])
    T.merge(pT, S, zs)
    for j in range(pT, me):
        T.step(j)
    bad = T.bad or Bf.bad
    y = Bf.ydot[1:] + T.ydot[1:]
    rdT, rdB = bandT[:, 0], bandB[:, 0]

    def back_T(x):                      # rows me-1 .. 0, in place on a D^-1-scaled vector
        for j in range(me - 1, -1, -1):
            for i in range(max(j - b, 0), j):
                x[i] -= bandT[i, j - i] * x[j]
        return x

    def back_B(x, xmid):                # xmid: middle values in global order
        known = xmid[::-1]
        for c in range(nB - 1, -1, -1):
            for o in range(1, b + 1):
                r = c + o
                if r >= nB + P:
                    continue
                x[c] -= bandB[c, o] * (x[r] if r < nB else known[r - nB])
        return x

    xT = back_T(zT[0] * rdT)
    xB = back_B(zB[0] * rdB, xT[pT:])
    u = np.concatenate([xT, xB[::-1]])
    # adjoint: forward-eliminated rhs = gy . (z_a0, z_a1) on bottom and middle, plus L_M^-1 w_mid
    wm = np.array(w_mid_fn(u[pT:me]), dtype=float)
    for j in range(pT, me):
        for i in range(j + 1, me):
            wm[i - pT] -= bandT[j, i - j] * wm[j - pT]
    pT_vec = np.zeros(me)
    pT_vec[pT:] = gy[0] * zT[1][pT:] + gy[1] * zT[2][pT:] + wm
    pB_vec = gy[0] * zB[1] + gy[1] * zB[2]
    xT2 = back_T(pT_vec * rdT)
    xB2 = back_B(pB_vec * rdB, xT2[pT:])
    psi = np.concatenate([xT2, xB2[::-1]])
    return u, y, psi, bad
