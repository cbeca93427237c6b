#!/usr/bin/env python
"""CPU model of alignw_reg_kernel<CPL> (csrc/pb_alignw.cu): the quality-weighted wavefront with the costs in registers, lane by
lane and step by step as the warp runs it -- slot ownership (lane L owns diagonals 2*CPL*L .. 2*CPL*(L+1)-1), the parity of the
steps, the ONE neighbour per step that comes from the adjacent lane, the two element windows that shift by one element on
alternate steps, the validity window of a step, the parents' layout, then goal cell and traceback from those parents.  Checked
against the extended oracle (pbo_align_weighted) by tests/test_wavefront_model.py; `lanes` < 32 models a group of a warp.
"""
W_INF = 0x3FFFFFFF
MATCH, INSERT, DELETE = 1, 2, 3


def derive(a_len, b_len, R):
    if b_len >= a_len:
        len_a = a_len
        D = 1 + int(len_a * R)
        len_b = min(b_len, len_a + D)
    else:
        len_b = b_len
        D = 1 + int(len_b * R)
        len_a = min(a_len, len_b + D)
    return len_a, len_b, D


def align_weighted_model(a, wa, b, wb, R, fail_scale, CPL, lanes=32, maxn=26000, maxm=6000):
    a_len, b_len = len(a), len(b)
    len_a, len_b, D = derive(a_len, b_len, R)
    out = dict(ret=-1, len_a=len_a, len_b=len_b, max_dst=D, matlen_a=0, matlen_b=0, cost=0, diag_cost=0, nedit=0, fail_row=0)
    if len_a >= maxn or D >= maxm:
        return out
    assert D + 1 <= lanes * CPL, "band too wide for this class"
    NS = 2 * CPL

    def elem(seq, w, x, n):  # outside a sequence: {0, weight 1}
        return (seq[x], w[x]) if 0 <= x < n else (0, 1)

    P1 = (1 + D) & 1
    c = [[0 if NS * L + s == D else W_INF for s in range(NS)] for L in range(lanes)]
    iTop, jBase, A, B, nA, nB = [], [], [], [], [], []
    for L in range(lanes):
        it1 = ((1 + D - P1) >> 1) - CPL * L
        it = it1 - (1 if P1 == 0 else 0)
        jb = 1 - it1 - (1 if P1 == 1 else 0)
        iTop.append(it); jBase.append(jb)
        A.append([elem(a, wa, it - 1 - u, len_a) for u in range(CPL)])
        B.append([elem(b, wb, jb - 1 + u, len_b) for u in range(CPL)])
        nA.append(elem(a, wa, it, len_a)); nB.append(elem(b, wb, jb + CPL - 1, len_b))
    nsteps, nfast = len_a + len_b, min(len_a, len_b)
    par = {}  # (d, lane, u) -> code
    laneD, sD = (D >> 1) // CPL, D - NS * ((D >> 1) // CPL)
    fail_row = 0
    for d in range(1, nsteps + 1):
        P = (d + D) & 1
        for L in range(lanes):  # window shift + the one new element
            if P == 0:
                A[L] = [nA[L]] + A[L][:-1]
                iTop[L] += 1
                nA[L] = elem(a, wa, iTop[L], len_a)
            else:
                B[L] = B[L][1:] + [nB[L]]
                jBase[L] += 1
                nB[L] = elem(b, wb, jBase[L] + CPL - 1, len_b)
            assert iTop[L] + jBase[L] == d
        i_lo, i_hi = max(0, d - len_b, (d - D + 1) >> 1), min(len_a, d)
        if P == 0:
            edge = [W_INF if L == 0 else c[L - 1][NS - 1] for L in range(lanes)]
        else:
            edge = [W_INF if L == lanes - 1 else c[L + 1][0] for L in range(lanes)]
        new = [row[:] for row in c]
        for L in range(lanes):
            for u in range(CPL):
                s = 2 * u + P
                i = iTop[L] - u
                left = edge[L] if s == 0 else c[L][s - 1]
                up = edge[L] if s == NS - 1 else c[L][s + 1]
                (ae, wai), (be, wbj) = A[L][u], B[L][u]
                cc = c[L][s] + (wai if ae != be else 0)
                code = MATCH
                t = left + wbj
                if t < cc:
                    cc, code = t, INSERT
                t = up + wai
                if t < cc:
                    cc, code = t, DELETE
                if i_lo <= i <= i_hi:
                    k = NS * L + s
                    assert 0 <= k <= 2 * D and i == (d + D - k) >> 1  # the cell the kernel means
                    new[L][s] = min(cc, W_INF)
                    par[(d, L, u)] = code
        c = new
        if d % 2 == 0:
            i = d >> 1
            if 10 < i <= nfast and float(c[laneD][sD]) > i * R * fail_scale:
                fail_row = i
                break
    out["fail_row"] = fail_row
    if fail_row:
        return out
    cst = {NS * L + s: c[L][s] for L in range(lanes) for s in range(NS)}
    # goal cell (seq_aligner.h:191-213): earliest strict minimum of the last row / column
    if len_a > len_b:
        cand = [(cst[len_b - i + D], i) for i in range(len_b, len_a + 1)]
    else:
        cand = [(cst[j - len_a + D], j) for j in range(len_a, len_b + 1)]
    best, pos = min(cand)
    matlen_a, matlen_b = (pos, len_b) if len_a > len_b else (len_a, pos)
    out.update(matlen_a=matlen_a, matlen_b=matlen_b, cost=best, diag_cost=cst[D] if (a_len <= len_a and a_len <= len_b) else 0)
    if float(matlen_b) < len_b * (1 - R):
        return out
    i, j, ops = matlen_a, matlen_b, []
    while (i > 0 or j > 0) and len(ops) < len_a + len_b + 1:
        k = j - i + D
        t = k >> 1
        code = par.get((i + j, t // CPL, t % CPL), 0)
        ops.append(code)
        if code == MATCH:
            i, j = i - 1, j - 1
        elif code == INSERT:
            j -= 1
        elif code == DELETE:
            i -= 1
        else:
            raise AssertionError("path left the computed cells")
    out.update(ret=matlen_b, nedit=len(ops), ops=ops[::-1])
    return out
