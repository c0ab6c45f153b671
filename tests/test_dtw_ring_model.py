"""CPU model of the index logistics of the DTW register wavefront (csrc/dtw.cu, dtw_fill_warp_body): which elements of
q / r the cp.async rings hold when the warp reads them.  The ring length and the refill period are read from the source,
so a change of either that lets a needed element be overwritten, or read while its copy may still be in flight, fails
here without a GPU (the band widths above 62 take the NPL = 8 instantiation, which no GPU test reaches)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "sonido-sonar_b200", "csrc", "dtw.cu")).read()
K_RING = int(re.search(r"constexpr int kWRing = (\d+);", SRC).group(1))
K_REFILL = int(re.search(r"constexpr int kWRefill = (\d+);", SRC).group(1))


def simulate(n, m, band, npl):
    H = npl // 2
    mask = K_RING - 1
    now_q, now_r = [None] * K_RING, [None] * K_RING      # element whose copy was issued last into the slot
    safe_q, safe_r = [None] * K_RING, [None] * K_RING    # ... as of the last cp.async.wait_group 0
    state = {"loaded": 0}

    def issue(target):
        for e in range(state["loaded"], target):
            if e < n:
                now_q[e & mask] = e
            if e < m:
                now_r[e & mask] = e
        state["loaded"] = max(state["loaded"], target)

    def refill(d):
        safe_q[:], safe_r[:] = now_q, now_r              # wait_group 0: everything issued so far has landed
        issue(((d + 2 * K_REFILL + band) >> 1) + 8)

    def rd(now, safe, idx):
        s = idx & mask
        return idx if (now[s] == idx and safe[s] == idx) else ("bad", idx, now[s], safe[s])

    lanes = range(32)
    geo = {}
    for lane in lanes:
        kbase = npl * lane - 1
        for x in range(npl):
            k = kbase + x
            delta = k - band
            lo = 2 + abs(delta)
            hi = min(2 * n - delta, 2 * m + delta)
            ok = 0 <= k <= 2 * band and hi >= lo
            geo[lane, x] = (lo, hi) if ok else None

    def valid(lane, x, d):
        g = geo[lane, x]
        return g is not None and g[0] <= d <= g[1]

    checked = 0

    def use(lane, x, d, qv, rv, i, j):
        nonlocal checked
        if not valid(lane, x, d):
            return
        assert 1 <= i <= n and 1 <= j <= m and abs(i - j) <= band and i + j == d, (lane, x, d, i, j)
        assert qv == i - 1, f"band {band} NPL {npl}: q[{i - 1}] expected at d={d} lane {lane}, register holds {qv}"
        assert rv == j - 1, f"band {band} NPL {npl}: r[{j - 1}] expected at d={d} lane {lane}, register holds {rv}"
        checked += 1

    d = 2
    issue(((d + K_REFILL + band) >> 1) + 8)
    refill(d)
    next_refill = d + K_REFILL
    Q = {lane: [None] * H for lane in lanes}
    R = {lane: [None] * (H + 1) for lane in lanes}
    if (d - band) & 1:
        for lane in lanes:
            ib = (d - 1 - band + npl * lane) >> 1
            jb = d - 1 - ib
            q = [rd(now_q, safe_q, ib + h - 1) for h in range(H)]
            r = [rd(now_r, safe_r, jb - u) for u in range(H + 1)]
            for h in range(H):  # half_a on diagonal d: even slots, cell (ib + h, jb - h + 1)
                use(lane, 2 * h, d, q[h], r[h], ib + h, jb - h + 1)
        d += 1
    ibs = {lane: (d - band + npl * lane) >> 1 for lane in lanes}
    for lane in lanes:
        ib, jb = ibs[lane], d - ibs[lane]
        Q[lane] = [rd(now_q, safe_q, ib + h - 1) for h in range(H)]
        R[lane] = [rd(now_r, safe_r, jb - u) for u in range(H + 1)]
    last = n + m
    while d <= last:
        if d >= next_refill:
            refill(d)
            next_refill += K_REFILL
        for lane in lanes:
            ib = ibs[lane]
            jb = d - ib
            qn, rn = rd(now_q, safe_q, ib + H - 1), rd(now_r, safe_r, jb + 1)
            for h in range(H):  # half_b: odd slots on d, cell (ib + h, jb - h): Q[h], R[h + 1]
                use(lane, 2 * h + 1, d, Q[lane][h], R[lane][h + 1], ib + h, jb - h)
            for h in range(H):  # half_a: even slots on d + 1, cell (ib + h, jb - h + 1): Q[h], R[h]
                use(lane, 2 * h, d + 1, Q[lane][h], R[lane][h], ib + h, jb - h + 1)
            Q[lane] = Q[lane][1:] + [qn]
            R[lane] = [rn] + R[lane][:-1]
            ibs[lane] = ib + 1
        d += 2
    return checked


def cells_in_band(n, m, band):
    return sum(1 for i in range(1, n + 1) for j in range(max(1, i - band), min(m, i + band) + 1))


@pytest.mark.parametrize("n,m,band,npl", [(700, 700, 50, 4), (900, 860, 5, 2), (650, 700, 30, 2), (800, 800, 62, 4),
                                          (800, 790, 63, 8), (900, 900, 100, 8), (1100, 1000, 126, 8), (40, 45, 20, 2)])
def test_every_cell_reads_its_own_elements(n, m, band, npl):
    assert 2 * band + 3 <= 32 * npl, "the instantiation must cover the band"
    assert K_REFILL + 126 + 10 <= K_RING, "the ring must span a refill period plus the widest band"
    assert simulate(n, m, band, npl) == cells_in_band(n, m, band)
