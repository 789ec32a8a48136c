"""Lane-level NumPy emulation of the warp-synchronous twisted banded LDL^T that
csrc/vbfem_band.cuh implements (test infrastructure).

It mirrors the kernel's index logic one to one -- 32 lanes, cyclic row
ownership, 26 rotating accumulator slots, gap reloads, mirrored bottom half,
middle merge, warp sweeps -- so that the slot/lane arithmetic can be validated
on a CPU against a dense solve before (and independently of) the CUDA code.
"""
import numpy as np

W = 32


class Half:
    """One elimination front in LOCAL coordinates: ``band[c, k] = A[c+k, c]`` for
    the ``ncols`` columns this front owns; rows run to ``nrows`` (> ncols when
    the trailing rows belong to the shared middle block)."""

    def __init__(self, band, ncols, nrows, b):
        self.band, self.ncols, self.nrows, self.b = band, ncols, nrows, b
        self.P = b + 1                         # slot period
        self.acc = np.zeros((W, self.P))
        self.zr = np.zeros(W)
        self.flag = False

    def _slot_col(self, R, s):
        return R - ((R - s) % self.P)

    def _load_slot(self, lanes, R, s):
        P = self.P
        for ln in lanes:
            r = int(R[ln])
            c = self._slot_col(r, s)
            ok = 0 <= c < self.ncols and r < self.nrows and r - c <= self.b
            self.acc[ln, s] = self.band[c, r - c] if ok else 0.0

    def init(self, z, j0=0):
        lane = np.arange(W)
        R = j0 + ((lane - j0) & 31)
        for s in range(self.P):
            self._load_slot(range(W), R, s)
        self.zr = np.array([z[r] if r < self.nrows else 0.0 for r in R])

    def step(self, j, z):
        """Eliminate local column j.  z: the front's rhs vector (read for new rows,
        written with the eliminated value z_j)."""
        b, P, band = self.b, self.P, self.band
        lane = np.arange(W)
        k = (lane - j) & 31
        R = j + k
        s0 = j % P
        live = (k <= b) & (R < self.nrows)
        v = np.where(live, self.acc[:, s0], 0.0)
        src = j & 31
        d, zj = v[src], self.zr[src]
        if not (d > 0.0 and np.isfinite(d)):
            self.flag = True
        rd = 1.0 / d
        w = v * rd
        # L column + inverse pivot go to the band; eliminated rhs value to z
        for ln in range(W):
            if 1 <= k[ln] <= b:
                band[j, k[ln]] = w[ln]
        band[j, 0] = rd
        z[j] = zj
        wv = band[j, :].copy()                 # broadcast read (13 x LDS.128)
        for m in range(1, b + 1):              # rank-1 update, garbage for m > k by design
            self.acc[:, (j + m) % P] -= v * wv[m]
        self.zr = np.where((k >= 1) & (k <= b), self.zr - w * zj, self.zr)
        # gap lanes (their row was eliminated 1..6 steps ago) reload slots for the next row
        gap = np.nonzero(k > b)[0]
        nper = -(-P // (W - P))                # slots per step so that 6 gap steps cover all 26
        for t in range(nper):
            self._load_slot(gap, R, (nper * j + t) % P)
        for ln in gap:
            self.zr[ln] = z[R[ln]] if R[ln] < self.nrows else 0.0

    def dump_middle(self, j, S, zS, n_mid):
        """After eliminating columns < j = ncols: write the Schur contributions of
        the rows >= ncols in MIRRORED (partner-front) band layout:
        S[a_lo, off] with a_lo = index of the lower partner row."""
        lane = np.arange(W)
        k = (lane - j) & 31
        for ln in range(W):
            if k[ln] >= n_mid:
                continue
            Rr = j + k[ln]                     # local row, middle index a = k
            for s in range(self.P):
                c = self._slot_col(Rr, s)
                if c >= j:                      # middle column q = c - j <= a
                    a, q = k[ln], c - j
                    S[n_mid - 1 - a, a - q] = self.acc[ln, s]
            zS[n_mid - 1 - k[ln]] = self.zr[ln]

    def merge_middle(self, j, S, zS, n_mid):
        lane = np.arange(W)
        k = (lane - j) & 31
        for ln in range(W):
            if k[ln] >= n_mid:
                continue
            Rr = j + k[ln]
            for s in range(self.P):
                c = self._slot_col(Rr, s)
                if c >= j:
                    self.acc[ln, s] += S[c - j, Rr - c]
            self.zr[ln] += zS[k[ln]]


def back_sweep(band, b, x, j_hi, j_lo, known=None):
    """In-place L^T x = y on local rows j_hi..j_lo (descending); x[j] for
    j > j_hi already final; ``known``: rows whose x is read, not produced."""
    lane = np.arange(W)
    r = j_hi - ((j_hi - lane) & 31)
    acc = np.array([x[q] if q >= 0 else 0.0 for q in r])
    for j in range(j_hi, j_lo - 1, -1):
        i = (j - lane) & 31
        xj = acc[j & 31]
        lv = np.array([band[r[ln], i[ln]] if (1 <= i[ln] <= b and r[ln] >= 0) else 0.0 for ln in range(W)])
        acc = acc - lv * xj
        src = j & 31
        x[j] = xj
        r[src] -= 32
        acc[src] = x[r[src]] if r[src] >= 0 else 0.0
    return x


def fwd_sweep(band, b, z, j_lo, j_hi, nrows):
    """In-place L z = w on local columns j_lo..j_hi (ascending), rows < nrows."""
    lane = np.arange(W)
    r = j_lo + ((lane - j_lo) & 31)
    acc = np.array([z[q] if q < nrows else 0.0 for q in r])
    for j in range(j_lo, j_hi + 1):
        k = (lane - j) & 31
        zj = acc[j & 31]
        lv = np.array([band[j, k[ln]] if (1 <= k[ln] <= b and r[ln] < nrows) else 0.0 for ln in range(W)])
        acc = acc - lv * zj
        src = j & 31
        z[j] = zj
        r[src] += 32
        acc[src] = z[r[src]] if r[src] < nrows else 0.0
    # rows beyond j_hi (the shared middle) keep their partial sums
    for ln in range(W):
        if j_hi < r[ln] < nrows:
            z[r[ln]] = acc[ln]
    return z


def twisted_solve(A, f, b, pT):
    """Solve A x = f (A SPD, half bandwidth b) with the twisted scheme: top front
    eliminates global columns [0, pT), bottom front (mirrored) eliminates
    [pT+b, n) downwards, the b x b middle is finished by the top front."""
    n = A.shape[0]
    nm = b
    nB = n - pT - nm
    LD = b + 1
    bandT = np.zeros((pT + nm, LD))
    bandB = np.zeros((max(nB, 1), LD))
    for hi in range(n):
        for lo in range(max(0, hi - b), hi + 1):
            if hi < pT + nm:
                bandT[lo, hi - lo] = A[hi, lo]
            else:
                bandB[n - 1 - hi, hi - lo] = A[hi, lo]
    zT = np.array(f[:pT + nm], dtype=float)
    zB = np.zeros(nB + nm)
    zB[:nB] = f[::-1][:nB]
    T, B = Half(bandT, pT + nm, pT + nm, b), Half(bandB, nB, nB + nm, b)
    T.init(zT)
    B.init(zB)
    for j in range(pT):
        T.step(j, zT)
    for j in range(nB):
        B.step(j, zB)
    S, zS = np.zeros((nm, LD)), np.zeros(nm)
    B.dump_middle(nB, S, zS, nm)
    T.merge_middle(pT, S, zS, nm)
    for j in range(pT, pT + nm):
        T.step(j, zT)
    # ---- back substitution: scale by D^-1, middle first, then both fronts
    xT = zT * bandT[:, 0]
    xB = np.zeros(nB + nm)
    xB[:nB] = zB[:nB] * bandB[:nB, 0]
    back_sweep(bandT, b, xT, pT + nm - 1, 0)   # (kernel: one continuous sweep; the bottom front starts
    xB[nB:] = xT[pT:][::-1]                     #  as soon as the middle rows are final)
    if nB > 0:
        back_sweep_known(bandB, b, xB, nB, nm)
    x = np.zeros(n)
    x[:pT + nm] = xT
    x[pT + nm:] = xB[:nB][::-1]
    return x, (bandT, bandB, T.flag or B.flag)


def back_sweep_known(band, b, x, ncols, nm):
    """Bottom front: rows ncols..ncols+nm-1 are known (middle, from the top
    front); they only feed the accumulators, then rows ncols-1..0 are solved."""
    lane = np.arange(W)
    j_hi = ncols + nm - 1
    r = (ncols - 1) - (((ncols - 1) - lane) & 31)
    acc = np.array([x[q] if q >= 0 else 0.0 for q in r])
    for j in range(j_hi, -1, -1):
        i = (j - lane) & 31
        xj = x[j] if j >= ncols else acc[j & 31]
        lv = np.array([band[r[ln], i[ln]] if (1 <= i[ln] <= b and 0 <= r[ln] < ncols and r[ln] + i[ln] == j)
                       else 0.0 for ln in range(W)])
        acc = acc - lv * xj
        if j < ncols:
            src = j & 31
            x[j] = xj
            r[src] -= 32
            acc[src] = x[r[src]] if r[src] >= 0 else 0.0
    return x
