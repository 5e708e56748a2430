"""NumPy emulation of the fused-kernel factorisation -- validates the
algebra the CUDA kernels rely on against the straightforward reference formulas:
  modulator:  x = M-IFFT_{m->n2} { Ctx[m][n1] * K-IFFT_{b->n1} { M-FFT(d_b)[m] } }
  receiver :  R_k[m] = K-FFT_{n1->k} { Crx[m][n1] * M-FFT_{n2->m}( x[n1 + K n2] ) },  y_k = M-IFFT(R_k)/M
with the pulse-shaping filter folded into the twiddle tables, and the two-pass padded
row-FFT addressing used in shared memory."""
import numpy as np

def ref_mod(d, taps, M, K, L):
    D = np.fft.fft(d.reshape(K, M), axis=1)
    X = np.zeros(M*K, complex); h = L//2; part = min(M*L//2, M)
    for k in range(K):
        for i in range(L):
            src = ((i+h) % L)*M; tgt = ((k+i+K-h) % K)*M
            X[tgt:tgt+part] += (D[k]*taps[src:src+M])[:part]
    return np.fft.ifft(X)

def ref_rx_fd(x, taps, M, K, L):
    Y = np.fft.fft(x); h = L//2; R = np.zeros((K, M), complex)
    for k in range(K):
        for i in range(L):
            src = ((k+i+K-h) % K)*M; tgt = ((i+h) % L)*M
            R[k] += taps[tgt:tgt+M]*Y[src:src+M]
    return R

def tx_table(taps, M, K, L):
    N = M*K; h = L//2; part = min(M*L//2, M)
    n1 = np.arange(K); C = np.zeros((M, K), complex)
    for m in range(M):
        if m >= part: continue
        G = sum(taps[((i+h) % L)*M+m]*np.exp(2j*np.pi*(i-h)*n1/K) for i in range(L))
        C[m] = G*np.exp(2j*np.pi*m*n1/N)/N
    return C

def rx_table(taps, M, K, L):
    N = M*K; h = L//2
    n1 = np.arange(K); C = np.zeros((M, K), complex)
    for m in range(M):
        G = sum(taps[((i+h) % L)*M+m]*np.exp(-2j*np.pi*(i-h)*n1/K) for i in range(L))
        C[m] = G*np.exp(-2j*np.pi*m*n1/N)
    return C

def row_fft_two_pass(row, R1, R2, sign):
    """in-place two-pass row FFT on the padded layout: element n at n + n//R2; result natural at [i]"""
    K = R1*R2
    buf = np.zeros(R1*(R2+1), complex)
    for n in range(K): buf[n + n//R2] = row[n]
    # pass 1: item n0: radix-R1 over n1 of x[R2*n1+n0], twiddle W_K^{n0 k1}
    for n0 in range(R2):
        v = np.array([buf[(R2+1)*n1+n0] for n1 in range(R1)])
        A = np.fft.fft(v) if sign < 0 else np.fft.ifft(v)*R1
        A = A*np.exp(sign*2j*np.pi*n0*np.arange(R1)/K)
        for k1 in range(R1): buf[(R2+1)*k1+n0] = A[k1]
    out = np.zeros(R1*(R2+1), complex)
    for k1 in range(R1):
        v = np.array([buf[(R2+1)*k1+n0] for n0 in range(R2)])
        X = np.fft.fft(v) if sign < 0 else np.fft.ifft(v)*R2
        for k0 in range(R2): out[k1+R1*k0] = X[k0]
    return out[:K]

def fused_mod(d, C, M, K, plan):
    D = np.fft.fft(d.reshape(K, M), axis=1)          # stage A: [k][m]
    rows = D.T.copy()                                 # [m][b]
    Z = np.stack([row_fft_two_pass(rows[m], *plan, +1) if plan[1] > 1 else np.fft.ifft(rows[m])*K for m in range(M)])
    Z = Z*C                                           # stage C
    x = np.fft.ifft(Z, axis=0)*M                      # M-IFFT over m -> [n2][n1]
    return x.reshape(M*K)                             # x[n1 + K n2]

def fused_rx(x, C, M, K, plan):
    U = np.fft.fft(x.reshape(M, K), axis=0)           # M-FFT over n2 -> [m][n1]
    U = U*C
    V = np.stack([row_fft_two_pass(U[m], *plan, -1) if plan[1] > 1 else np.fft.fft(U[m]) for m in range(M)])
    return V.T                                        # R[k][m]

rng = np.random.default_rng(0)
for (M, K, L, plan) in [(5,16,2,(16,1)), (9,64,2,(8,8)), (15,256,2,(16,16)), (15,1024,2,(32,32)), (15,512,2,(32,16)),
                        (7,128,4,(16,8)), (6,64,1,(8,8)), (15,128,3,(16,8))]:
    taps = rng.standard_normal(M*L)+1j*rng.standard_normal(M*L)
    d = rng.standard_normal(M*K)+1j*rng.standard_normal(M*K)
    e1 = np.abs(fused_mod(d, tx_table(taps,M,K,L), M, K, plan)-ref_mod(d,taps,M,K,L)).max()
    e2 = np.abs(fused_rx(d, rx_table(taps,M,K,L), M, K, plan)-ref_rx_fd(d,taps,M,K,L)).max() if L >= 2 else 0
    print((M,K,L,plan), 'mod err %.2e  rx err %.2e' % (e1, e2))


# ---------------------------------------------------------------------------------------------
# v3 kernels: commuted form, row FFTs IN PLACE on the [k][t]-ordered frame, XOR-swizzled exchange.
#   modulator: S[k*M+t] = d           -> rows t: K-IFFT over k (in place)  -> column n1: FFT_M, *Ctx, IFFT_M -> x
#   receiver : column n1: FFT_M(x), *Crx, IFFT_M/M -> S[n1*M+t] -> rows t: K-FFT over n1 (in place) -> S = y[k][t]
def row_fft_inplace_xor(S, M, t, R, sign, single):
    """row t of the frame held as S[q*M + t], q = 0..K-1; natural order in and out."""
    K = R if single else R * R
    if single:
        v = np.array([S[q * M + t] for q in range(K)])
        X = np.fft.fft(v) if sign < 0 else np.fft.ifft(v) * K
        for q in range(K): S[q * M + t] = X[q]
        return
    regs = {}
    for n0 in range(R):   # pass 1, lane n0: registers a[i] = x[R*i + n0]
        a = np.array([S[(R * i + n0) * M + t] for i in range(R)])
        A = np.fft.fft(a) if sign < 0 else np.fft.ifft(a) * R
        regs[n0] = A * np.exp(sign * 2j * np.pi * n0 * np.arange(R) / K)
    for n0 in range(R):   # exchange write (after every lane has read): slot R*k1 + (n0 ^ k1)
        for k1 in range(R): S[(R * k1 + (n0 ^ k1)) * M + t] = regs[n0][k1]
    regs2 = {}
    for k1 in range(R):   # pass 2, lane k1: registers b[n0]
        b = np.array([S[(R * k1 + (n0 ^ k1)) * M + t] for n0 in range(R)])
        regs2[k1] = np.fft.fft(b) if sign < 0 else np.fft.ifft(b) * R
    for k1 in range(R):   # natural-order write (after every lane has read)
        for k0 in range(R): S[(k1 + R * k0) * M + t] = regs2[k1][k0]


def v3_mod(d, C, M, K, R, single):
    S = d.astype(complex).copy()
    for t in range(M): row_fft_inplace_xor(S, M, t, R, +1, single)
    x = np.zeros(M * K, complex)
    for n1 in range(K):
        v = np.fft.fft(S[n1 * M:(n1 + 1) * M]) * C[:, n1]
        x[n1 + K * np.arange(M)] = np.fft.ifft(v) * M
    return x


def v3_rx(x, C, M, K, R, single, td=True):
    S = np.zeros(M * K, complex)
    for n1 in range(K):
        v = np.fft.fft(x[n1 + K * np.arange(M)]) * C[:, n1]
        S[n1 * M:(n1 + 1) * M] = np.fft.ifft(v) if td else v
    for t in range(M): row_fft_inplace_xor(S, M, t, R, -1, single)
    return S.reshape(K, M)


print('v3 (commuted, in place):')
for (M, K, L, R, single) in [(5, 16, 2, 16, True), (9, 64, 2, 8, False), (15, 256, 2, 16, False), (15, 1024, 2, 32, False)]:
    taps = rng.standard_normal(M * L) + 1j * rng.standard_normal(M * L)
    d = rng.standard_normal(M * K) + 1j * rng.standard_normal(M * K)
    e1 = np.abs(v3_mod(d, tx_table(taps, M, K, L), M, K, R, single) - ref_mod(d, taps, M, K, L)).max()
    Rfd = ref_rx_fd(d, taps, M, K, L)
    e2 = np.abs(v3_rx(d, rx_table(taps, M, K, L), M, K, R, single, td=False) - Rfd).max()
    e3 = np.abs(v3_rx(d, rx_table(taps, M, K, L), M, K, R, single, td=True) - np.fft.ifft(Rfd, axis=1)).max()
    print((M, K, L, R), 'mod err %.2e  rx fd err %.2e  rx td err %.2e' % (e1, e2, e3))


# ---------------------------------------------------------------------------------------------
# two-pass kernels (K = 2*K1, frame larger than shared memory): decimation in frequency over the
# subcarrier index -- pass p computes the K1-point transform that yields the outputs of parity p.
#   modulator: E^p_b' = (d_b' + (-1)^p d_{b'+K1}) W^{p b'} (raw symbols, W = e^{+j2pi/K})
#              Z_m[2n'+p] = K1-IFFT_{b'->n'} FFT_M(E^p_b')[m];  x[(2n'+p) + K n2] = M-IFFT_m Ctx[m][2n'+p] Z_m[2n'+p]
#   receiver : B^p_m[n'] = Tlo^p[m][n'] U_n'[m] + Thi^p[m][n'] U_{n'+K1}[m],
#              Tlo^p = Crx[m][n'] W^{p n'}, Thi^p = (-1)^p Crx[m][n'+K1] W^{p n'}  (W = e^{-j2pi/K})
#              R_{2k'+p}[m] = K1-FFT_{n'->k'} B^p_m[n']
def twopass_mod(d, C, M, K):
    K1 = K // 2
    dd = d.reshape(K, M)
    x = np.zeros(M * K, complex)
    for p in range(2):
        E = (dd[:K1] + (-1) ** p * dd[K1:]) * np.exp(2j * np.pi * p * np.arange(K1) / K)[:, None]
        D = np.fft.fft(E, axis=1)                         # [b'][m]
        Z = np.fft.ifft(D.T, axis=1) * K1                 # [m][n']
        Z = Z * C[:, p::2]
        xp = np.fft.ifft(Z, axis=0) * M                   # [n2][n']
        for n2 in range(M): x[p + 2 * np.arange(K1) + K * n2] = xp[n2]
    return x


def twopass_rx(x, C, M, K):
    K1 = K // 2
    U = np.fft.fft(x.reshape(M, K), axis=0)               # [m][n1]
    R = np.zeros((K, M), complex)
    for p in range(2):
        w = np.exp(-2j * np.pi * p * np.arange(K1) / K)
        B = C[:, :K1] * w * U[:, :K1] + (-1) ** p * C[:, K1:] * w * U[:, K1:]
        V = np.fft.fft(B, axis=1)                         # [m][k']
        R[p::2] = V.T
    return R


print('two-pass (K = 2*K1, decimation in frequency over subcarriers):')
for (M, K, L) in [(15, 2048, 2), (15, 64, 2), (9, 128, 3)]:
    taps = rng.standard_normal(M * L) + 1j * rng.standard_normal(M * L)
    d = rng.standard_normal(M * K) + 1j * rng.standard_normal(M * K)
    e1 = np.abs(twopass_mod(d, tx_table(taps, M, K, L), M, K) - ref_mod(d, taps, M, K, L)).max()
    e2 = np.abs(twopass_rx(d, rx_table(taps, M, K, L), M, K) - ref_rx_fd(d, taps, M, K, L)).max()
    print((M, K, L), 'mod err %.2e  rx fd err %.2e' % (e1, e2))
