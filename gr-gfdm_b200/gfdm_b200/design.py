"""Host-side design utilities (NumPy, Python 3): everything a caller needs to
build the *constructor inputs* of the kernels -- prototype-filter taps, receive
taps (MF / ZF), ramp windows, subcarrier maps, preambles, constellations.

These are configuration-time computations (run once per handle), not the data
path.  They restate the reference's pygfdm helpers without its commpy / NumPy<1.24
dependencies and are validated against pygfdm-generated golden vectors in
tests/test_design.py:
  python/pygfdm/filters.py:27-54        get_frequency_domain_filter
  python/pygfdm/cyclic_prefix.py:35-68  windows
  python/pygfdm/mapping.py:53-81        resource maps
  python/pygfdm/zadoff_chu.py:11-24     Zadoff-Chu sequence
  python/pygfdm/preamble.py:91-132      mapped_preamble
  python/pygfdm/utils.py:37-44          get_random_qpsk
  python/pygfdm/configurator.py:39-82   get_gfdm_configuration
The ZF receive taps follow python/gfdmlib/gfdm/detail/{gfdmutil.py:107-131,
gabor.py,Demodulator.py:46-51} (Gabor dual window, sampled to the sparse
frequency-domain taps); the reference never exercises them through its C++
kernels, see DESIGN.md "parity unpinned".
"""
from collections import namedtuple

import numpy as np


# --------------------------------------------------------------------------
# prototype filters
def rrc_impulse(n, alpha, ts, fs=1.0):
    """Root-raised-cosine impulse response, n samples centred on n/2
    (same closed form and sampling grid as commpy.rrcosfilter)."""
    n = int(n)
    t = (np.arange(n) - n / 2) / float(fs)
    h = np.empty(n, dtype=float)
    with np.errstate(divide='ignore', invalid='ignore'):
        num = np.sin(np.pi * t * (1 - alpha) / ts) + 4 * alpha * (t / ts) * np.cos(np.pi * t * (1 + alpha) / ts)
        den = np.pi * t * (1 - (4 * alpha * t / ts) ** 2) / ts
        h[:] = num / den
    h[t == 0.0] = 1.0 - alpha + (4 * alpha / np.pi)
    if alpha != 0:
        sing = (t == ts / (4 * alpha)) | (t == -ts / (4 * alpha))
        h[sing] = (alpha / np.sqrt(2)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) +
                                          (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
    return h


def rc_impulse(n, alpha, ts, fs=1.0):
    """Raised-cosine impulse response (commpy.rcosfilter closed form)."""
    n = int(n)
    t = (np.arange(n) - n / 2) / float(fs)
    h = np.empty(n, dtype=float)
    with np.errstate(divide='ignore', invalid='ignore'):
        h[:] = (np.sin(np.pi * t / ts) / (np.pi * t / ts)) * \
               (np.cos(np.pi * alpha * t / ts) / (1 - ((2 * alpha * t) / ts) ** 2))
    h[t == 0.0] = 1.0
    if alpha != 0:
        sing = (t == ts / (2 * alpha)) | (t == -ts / (2 * alpha))
        with np.errstate(divide='ignore', invalid='ignore'):
            h[sing] = (np.pi / 4) * (np.sin(np.pi * t[sing] / ts) / (np.pi * t[sing] / ts))
    return h


def gfdm_filter_taps(filtertype, alpha, M, K, oversampling_factor=1):
    n = M * K * oversampling_factor
    if filtertype == 'rrc':
        return rrc_impulse(n, alpha, 1. * K * oversampling_factor, 1.)
    if filtertype == 'rc':
        return rc_impulse(n, alpha, 1. * K * oversampling_factor, 1.)
    raise ValueError('unknown filtertype %r' % (filtertype,))


def gfdm_freq_taps(h):
    return np.fft.fft(np.roll(h, h.shape[-1] // 2))


def gfdm_freq_taps_sparse(H, M, L):
    # `H[-(M * L) // 2:]` as the reference writes it (python/pygfdm/filters.py:43): floor(-(M*L)/2), i.e. for odd M*L
    # the negative half is one tap longer than the positive half and the result has M*L entries
    return np.concatenate((H[0:(M * L) // 2], H[-(M * L) // 2:]))


def get_frequency_domain_filter(filtertype, alpha, M, K, L):
    """L*M frequency-domain taps in FFT order, scaled so that H.H = M."""
    H = gfdm_freq_taps_sparse(gfdm_freq_taps(gfdm_filter_taps(filtertype, alpha, M, K, 1)), M, L)
    return H / np.sqrt(H.dot(H).real / M)


def get_matched_filter_taps(tx_taps):
    """MF receive taps (python/pygfdm/configurator.py:79)."""
    return np.conjugate(tx_taps)


def get_zero_forcing_taps(filtertype, alpha, M, K, L=2):
    """"ZF" receive taps for the sparse receiver (recipe of gfdmlib's ZFFFT demodulator,
    python/gfdmlib/gfdm/detail/Demodulator.py:24-27 with gfdmutil.py:111-112):
    canonical Gabor dual of the transmit pulse on the critically sampled (K, M)
    lattice -- Zak domain: Z(gamma) = 1 / conj(Z(g)) up to scale -- subsampled in
    time by K/L, i.e. its spectrum folded onto the L*M sparse taps, conjugated for
    the correlation receiver and renormalised like every tap set (the receiver
    kernel renormalises again, so only the shape matters).  With L = 2 the fold
    ALIASES the dual's wide spectrum onto two blocks: measured through the reference
    kernels this is a worse receiver than the matched filter (tests/test_design.py),
    it is kept because SURVEY 8d prescribes it for the benchmark taps.  Use
    get_zero_forcing_taps_full for a receiver that zero-forces.
    """
    if K % L:
        raise ValueError('subcarriers MUST be a multiple of overlap')
    N = M * K
    g = np.roll(gfdm_filter_taps(filtertype, alpha, M, K, 1), N // 2)  # pulse peak at index 0
    Zg = np.fft.fft(g.reshape(M, K), axis=0)    # Zak transform: [m, k] <- g[k + l K]
    Zd = 1.0 / np.conjugate(Zg)
    gamma = np.fft.ifft(Zd, axis=0).reshape(N)  # dual window
    G = np.conjugate(np.fft.fft(gamma[::K // L]))
    return G / np.sqrt(np.sum(np.abs(G) ** 2) / M)


def pulse_spectrum_from_sparse_taps(taps, M, K, L):
    """N-bin spectrum (FFT order) of the pulse the kernels implement with L*M sparse taps: the modulator scatters
    taps[((i+h)%L)*M + m] onto bin ((k+i-h) mod K)*M + m (lib/modulator_kernel_cc.cc:113-135), h = L/2, so subcarrier 0's
    pulse has G[((i-h) mod K)*M + m] = taps[((i+h)%L)*M + m]; every other bin is zero.  Taps are renormalised like the
    kernels do (sum |T|^2 = M)."""
    taps = np.asarray(taps, dtype=complex)
    taps = taps / np.sqrt(np.sum(np.abs(taps) ** 2) / M)
    G = np.zeros(M * K, dtype=complex)
    h = L // 2
    for i in range(L):
        b = (i - h) % K
        G[b * M:(b + 1) * M] += taps[((i + h) % L) * M:((i + h) % L) * M + M]
    return G


def sparse_taps_from_pulse_spectrum(G, M, K, L):
    """Inverse of pulse_spectrum_from_sparse_taps: the L*M taps (kernel tap order) that keep the L blocks of M bins
    around DC of an N-bin spectrum; L = K keeps everything."""
    G = np.asarray(G)
    h = L // 2
    T = np.zeros(L * M, dtype=complex)
    for i in range(L):
        b = (i - h) % K
        T[((i + h) % L) * M:((i + h) % L) * M + M] = G[b * M:(b + 1) * M]
    return T


def _zak(g, M, K):
    return np.fft.fft(np.reshape(g, (M, K)), axis=0)  # [m, k] <- sum_l g[k + l K] e^{-j 2 pi m l / M}


def get_zero_forcing_taps_full(tx_taps, M, K, L, rx_overlap=None):
    """Receive taps that ZERO-FORCE the modulator built from `tx_taps` (L*M sparse taps): the canonical dual window of
    the transmit pulse on the critically sampled (K, M) Gabor lattice, Z(gamma) = 1 / conj(Z(g)) in the Zak domain
    (python/gfdmlib/gfdm/detail/gfdmutil.py:107-113, gabor.py:41-72), expressed as receive taps of
    receiver_kernel_cc with overlap = K (the whole spectrum, FFT order; default) -- noiseless demodulation then returns
    the transmitted symbols times a real gain (the receiver renormalises its taps).  With rx_overlap < K the dual's
    spectrum is truncated to the rx_overlap blocks around DC: an approximation whose residual interference falls
    slowly for wide pulses (RRC 0.5, M = 15: EVM 0.36 / 0.07 / 0.013 at overlap 2 / 16 / 32).  Requires an invertible
    transmit matrix (odd M for the usual symmetric pulses)."""
    G = pulse_spectrum_from_sparse_taps(tx_taps, M, K, L)
    g = np.fft.ifft(G)                       # x[n] = sum_{k,t} d[k,t] g[n - tK] e^{j 2 pi k n / K}
    Zg = _zak(g, M, K)
    if np.min(np.abs(Zg)) < 1e-9 * np.max(np.abs(Zg)):
        raise ValueError('the transmit matrix is singular (zero of the Zak transform): no zero-forcing receiver exists')
    gamma = np.fft.ifft(1.0 / np.conjugate(Zg), axis=0).reshape(M * K)
    Gam = np.conjugate(np.fft.fft(gamma))
    Lr = K if rx_overlap is None else int(rx_overlap)
    T = sparse_taps_from_pulse_spectrum(Gam, M, K, Lr)
    return T / np.sqrt(np.sum(np.abs(T) ** 2) / M)


def get_mmse_taps_full(tx_taps, M, K, L, snr_db=20.0, rx_overlap=None):
    """Linear MMSE receive taps for the modulator built from `tx_taps`: B = (A^H A + sigma^2 I)^-1 A^H is diagonal in
    the Zak domain on the critical lattice (eigenvalues of A^H A are K |Z(g)|^2), so
    Z(gamma) = Z(g) / (|Z(g)|^2 + sigma^2 / K) (gfdmlib's MMSE pulse, gfdmutil.py:115-131, written without its masking
    of small values).  sigma^2 = mean transmit sample power / 10^(snr_db/10) for unit-energy symbols.  snr -> inf is
    get_zero_forcing_taps_full, snr -> -inf the matched filter."""
    G = pulse_spectrum_from_sparse_taps(tx_taps, M, K, L)
    g = np.fft.ifft(G)
    sigma2 = np.sum(np.abs(g) ** 2) * 10.0 ** (-snr_db / 10.0)
    Zg = _zak(g, M, K)
    gamma = np.fft.ifft(Zg / (np.abs(Zg) ** 2 + sigma2 / K), axis=0).reshape(M * K)
    Gam = np.conjugate(np.fft.fft(gamma))
    Lr = K if rx_overlap is None else int(rx_overlap)
    T = sparse_taps_from_pulse_spectrum(Gam, M, K, Lr)
    return T / np.sqrt(np.sum(np.abs(T) ** 2) / M)


def get_mmse_taps(filtertype, alpha, M, K, L=2, snr_db=20.0):
    """MMSE receive taps for the sparse receiver: gfdmlib's MMSE receiver pulse
    (python/gfdmlib/gfdm/detail/gfdmutil.py:115-131, Zak transforms of gabor.py:11-21) restated for NumPy/Py3 --
    Z(g_mmse) = 1 / conj(Z(g) + sigma^2 Z(g_zf) / K), zeros of Z(g) masked (|Z| < 1e-6), sigma^2 = 10^(-snr/10) --
    then folded onto the L*M sparse taps exactly like get_zero_forcing_taps.  snr -> inf gives the ZF taps,
    snr -> -inf the matched filter.  Design parity: unpinned (gfdmlib is Python 2 and is not exercised by the
    reference's kernels); the kernels' arithmetic with given taps is pinned.
    """
    if K % L:
        raise ValueError('subcarriers MUST be a multiple of overlap')
    N = M * K
    sigma2 = 10.0 ** (-snr_db / 10.0)
    g = np.roll(gfdm_filter_taps(filtertype, alpha, M, K, 1), N // 2)  # pulse peak at index 0
    Zg = np.fft.fft(g.reshape(M, K), axis=0)                            # sqrt(K) * zak(g, K), as [m, k]
    small = np.abs(Zg) < 1e-6
    Zg2 = np.where(small, 1.0, Zg)
    Zgd = 1.0 / np.conjugate(Zg2)
    Zgm = 1.0 / np.conjugate(Zg + sigma2 * Zgd / K)
    Zgm[small] = 0.0
    gmm = np.fft.ifft(Zgm, axis=0).reshape(N)
    gmm = gmm / np.sum(gmm * g)
    G = np.conjugate(np.fft.fft(gmm[::K // L]))
    return G / np.sqrt(np.sum(np.abs(G) ** 2) / M)


# --------------------------------------------------------------------------
# windows
def get_window_len(cp_len, n_timeslots, n_subcarriers, cs_len=0):
    return n_timeslots * n_subcarriers + cp_len + cs_len


def window_ramp(ramp_len, window_len):
    r = np.array([]) if ramp_len < 1 else np.arange(0, 1, 1. / ramp_len)
    return np.concatenate((1. - r, np.zeros(window_len - 2 * ramp_len), r))


def get_raised_cosine_ramp(ramp_len, window_len):
    return .5 * (1. + np.cos(np.pi * window_ramp(ramp_len, window_len)))


def get_root_raised_cosine_ramp(ramp_len, window_len):
    return np.sqrt(get_raised_cosine_ramp(ramp_len, window_len))


# --------------------------------------------------------------------------
# mapping
def get_subcarrier_map(subcarriers, active_subcarriers, dc_free=False):
    if dc_free:
        return np.concatenate((np.arange(1, active_subcarriers // 2 + 1),
                               np.arange(subcarriers - active_subcarriers // 2, subcarriers)))
    return np.concatenate((np.arange(0, active_subcarriers // 2),
                           np.arange(subcarriers - active_subcarriers // 2, subcarriers)))


def default_active_subcarriers(subcarriers):
    """A = 2 * floor(0.8125 * K / 2) (rule of python/qa_vc_compatibility_check.py:160-161)."""
    return 2 * int(0.8125 * subcarriers / 2)


# --------------------------------------------------------------------------
# symbols
def get_random_qpsk(nsamples, seed=None, rng=None):
    """(+-1 +-1j)/sqrt(2); with `seed` reproduces pygfdm.utils.get_random_qpsk(n, seed)."""
    if rng is None:
        rng = np.random.RandomState(seed) if seed else np.random.RandomState()
    d = rng.randint(0, 2, 2 * nsamples) * -2. + 1.
    d = np.reshape(d, (2, -1))
    return (d[0] + 1j * d[1]) / np.sqrt(2)


def qam16_points():
    lv = np.array([-3., -1., 1., 3.])
    return (lv[:, None] + 1j * lv[None, :]).flatten() / np.sqrt(10.)


def get_random_qam16(nsamples, rng):
    return qam16_points()[rng.integers(0, 16, nsamples)]


def generate_zadoff_chu_sequence(seq_length, uvalue, shift_value=0):
    if not np.gcd(seq_length, uvalue) == 1:
        raise RuntimeError('GCD(N_ZC=%d, u=%d) != 1 !' % (seq_length, uvalue))
    if not 0 < uvalue < seq_length:
        raise RuntimeError('Does not satisfy: 0 < u=%d < N_ZC!' % uvalue)
    c_f = seq_length % 2
    n = np.arange(seq_length)
    return np.exp(-1.j * np.pi * (n * (n + c_f + 2 * shift_value)) / seq_length)


# --------------------------------------------------------------------------
# preamble (needs a tiny M=2 modulation; NumPy spec, configuration time only)
def _modulate_block_spec(d_kmajor, taps, M, K, L):
    """NumPy statement of the sparse-frequency-domain modulator for ONE block,
    used only to synthesise the (M=2) preamble at configuration time."""
    N = M * K
    D = np.fft.fft(np.reshape(d_kmajor, (K, M)), axis=1)
    X = np.zeros(N, dtype=complex)
    part = min(M * L // 2, M)
    for k in range(K):
        for i in range(L):
            src = ((i + L // 2) % L) * M
            tgt = ((k + i + K - L // 2) % K) * M
            X[tgt:tgt + part] += (D[k] * taps[src:src + M])[:part]
    return np.fft.ifft(X)


def map_to_waveform_resources(syms, active_subcarriers, fft_len, subcarrier_map, per_timeslot=True):
    n = len(syms)
    ts = int(np.ceil(1. * n / active_subcarriers))
    s = np.concatenate((syms, np.zeros(active_subcarriers * ts - n)))
    s = np.reshape(s, [-1, active_subcarriers]).T if per_timeslot else np.reshape(s, [active_subcarriers, -1])
    frame = np.zeros((fft_len, ts), dtype=np.complex64)  # complex64 grid, as python/pygfdm/mapping.py:73
    frame[subcarrier_map, :] = s
    return frame.flatten()


def mapped_preamble(seed, filtertype, alpha, active_subcarriers, fft_len, subcarrier_map, overlap,
                    cp_len, ramp_len, use_zadoff_chu=False, cyclic_shift=0):
    """Returns (full preamble with CP/CS/window, core 2*fft_len preamble)."""
    if use_zadoff_chu:
        pn = generate_zadoff_chu_sequence(active_subcarriers, 19)
    else:
        pn = get_random_qpsk(active_subcarriers, seed)
    pn_sym = map_to_waveform_resources(pn, active_subcarriers, fft_len, subcarrier_map)
    M = 2
    H = get_frequency_domain_filter(filtertype, alpha, M, fft_len, overlap)
    H = H * (1. / np.sqrt(np.sum(np.abs(H) ** 2) / 2))
    # both timeslots carry the same symbols: d[k*2 + m] = pn_sym[k]
    core = _modulate_block_spec(np.repeat(pn_sym, M), H, M, fft_len, overlap)
    sym = np.concatenate((core[-cp_len:], core, core[0:ramp_len]))
    sym = np.roll(sym, cyclic_shift)
    sym = sym * get_raised_cosine_ramp(ramp_len, get_window_len(cp_len, M, fft_len, ramp_len))
    return sym, core


PREAMBLE_SEED = int(3660365253)

GfdmConfiguration = namedtuple('GfdmConfiguration', [
    'timeslots', 'subcarriers', 'active_subcarriers', 'overlap', 'cp_len', 'cs_len', 'ramp_len',
    'cyclic_shifts', 'seed', 'block_len', 'window_len', 'subcarrier_map', 'full_preambles',
    'core_preamble', 'preamble_len', 'core_preamble_len', 'frame_len', 'window_taps',
    'tx_filter_taps', 'rx_filter_taps'])


def get_gfdm_configuration(timeslots=9, subcarriers=64, active_subcarriers=52, overlap=2, cp_len=16,
                           cs_len=8, filtertype='rrc', filteralpha=0.2, cyclic_shifts=(0,)):
    """Bundle of constructor inputs (python/pygfdm/configurator.py:39-82, minus padding)."""
    ramp_len = cs_len
    block_len = timeslots * subcarriers
    window_len = block_len + cp_len + cs_len
    smap = get_subcarrier_map(subcarriers, active_subcarriers, dc_free=True)
    pre = [mapped_preamble(PREAMBLE_SEED, filtertype, filteralpha, active_subcarriers, subcarriers,
                           smap, overlap, cp_len, ramp_len, use_zadoff_chu=True, cyclic_shift=cs)
           for cs in cyclic_shifts]
    full = [p[0] for p in pre]
    core = pre[0][1]
    tx = get_frequency_domain_filter(filtertype, filteralpha, timeslots, subcarriers, overlap)
    return GfdmConfiguration(
        timeslots, subcarriers, active_subcarriers, overlap, cp_len, cs_len, ramp_len,
        list(cyclic_shifts), PREAMBLE_SEED, block_len, window_len, smap, full, core, full[0].size,
        core.size, window_len + full[0].size,
        get_raised_cosine_ramp(ramp_len, get_window_len(cp_len, timeslots, subcarriers, cs_len)),
        tx, np.conjugate(tx))
