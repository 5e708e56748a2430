#!/usr/bin/env python3
"""Pick (R1, R2, T, IPT, MINB) for the fused single-pass kernels of a (M, K) shape: mirrors the constexpr arithmetic of
`Shape` in gr-gfdm_b200/csrc/fused_dev.cuh (shared-memory budget, tensor-memory columns, static_asserts) and a register
estimate, and prints the GFDM_SHAPE lines of csrc/fused_shapes_*.cu.  usage: python tools/shape_chooser.py"""
import itertools


def shape(M, R1, R2, T, IPT, MINB, tmem=True):
    K = R1 * R2
    N = M * K
    if IPT * T % K or IPT * T // K < 1 or 32 % R1:
        return None
    if not (R2 == 1 or (R2 & (R2 - 1) == 0 and R2 <= 32)):
        return None
    F = IPT * T // K
    two = R2 > 1
    RS = R1 * (R2 + 1) if two else (K | 1)
    ROWS = F * M
    BUF = max(ROWS * RS, F * N)
    BUDGET = ((233472 // MINB) - 1024) // 8 - 8 - 192
    tbl_smem = BUDGET - BUF - (K if two else 0) >= F * N + N
    tbl_tmem = (not tbl_smem) and tmem
    tw_tmem = tbl_tmem and two and T % R2 == 0
    TW = K if (two and not tw_tmem) else 0
    P_MAX = BUDGET - BUF - TW
    if P_MAX < 0:
        return None
    TBL = N if tbl_smem else 0
    P_AVAIL = P_MAX - TBL
    PF = min(F * N, (P_AVAIL // (2 * M)) * 2 * M)
    PR = min(M, P_AVAIL // (F * K))
    if PR < 1 or PF < 2 * M:
        return None
    per_thread = IPT * 2 * M + (2 * R1 if tw_tmem else 0)
    used = ((T // 32 + 3) // 4) * per_thread
    cols = next(c for c in (32, 64, 128, 256, 512, 1 << 20) if used <= c)
    if tbl_tmem and (used > 512 or cols * MINB > 512):
        return None
    if not tbl_smem and not tbl_tmem:
        return None
    if tw_tmem and R1 % 8:
        return None
    if T * MINB > 2048:
        return None
    regs_avail = min(255, 65536 // (T * MINB))
    # stage A/C keep IPT*M complex values live (+ the M-point transform's temporaries), the row passes 2*radix (+ twiddles)
    # (+ margin: the transmitter-chain and cancellation variants of the same shape carry ~20 more live values)
    regs_need = max(2 * M * IPT + 34, 2 * max(R1, R2 if two else 0) + 40) + (24 if IPT == 1 else 12)
    if regs_need > regs_avail:
        return None
    # one subcarrier per thread first (the cancellation loop needs it), then resident threads, then prefetch depth
    score = (IPT == 1, MINB * T, PR)
    return dict(M=M, R1=R1, R2=R2, T=T, IPT=IPT, MINB=MINB, F=F, tbl='smem' if tbl_smem else 'tmem', PR=PR, PF=PF,
                regs_avail=regs_avail, regs_need=regs_need, score=score, smem=8 * (BUF + TW + TBL + max(PF, F * PR * K) + 192) + 64)


def choose(M, K):
    best = None
    for R1 in (4, 8, 16, 32):
        if K % R1:
            continue
        R2 = K // R1
        if R2 > 32 or (R2 > 1 and R1 > R2 * 4):
            continue
        for T, IPT, MINB in itertools.product((128, 256, 512), (1, 2), (1, 2, 3, 4)):
            s = shape(M, R1, R2, T, IPT, MINB)
            if s and (best is None or s['score'] > best['score'] or (s['score'] == best['score'] and abs(R1 - max(R2, 1)) < abs(best['R1'] - max(best['R2'], 1)))):
                best = s
    return best


if __name__ == '__main__':
    existing = {(5, 16), (9, 64), (15, 256), (15, 1024)}
    shapes = [(M, K) for K in (16, 32, 64, 128, 256, 512, 1024) for M in (3, 5, 7, 9, 15, 21)] + [(8, 16), (16, 4), (7, 8), (19, 32), (16, 64)]
    for M, K in shapes:
        if (M, K) in existing:
            continue
        s = choose(M, K)
        if s is None:
            print('// (M=%d, K=%d): no single-pass configuration fits' % (M, K))
            continue
        print('GFDM_SHAPE(%d, %d, %d, %d, %d, %d), // K=%d: %d frame(s) per pass, table in %s, PR=%d, %d B smem, regs %d/%d'
              % (s['M'], s['R1'], s['R2'], s['T'], s['IPT'], s['MINB'], K, s['F'], s['tbl'], s['PR'], s['smem'], s['regs_need'], s['regs_avail']))
