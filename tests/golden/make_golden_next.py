#!/usr/bin/env python3
"""Golden vectors for the rows either side of the hot path:
symbol mapping, remove_prefix_cc and extract_burst_cc -> tests/golden/next_rows.npz.

Runs ONLY in the build container (needs /root/reference).  Sources of the expected values:
 * symbol mapping: UNMODIFIED pygfdm code (python/pygfdm/symbolmapping.py bits2symbols / symbols2bits /
   pack_bits, python/pygfdm/utils.py get_random_qpsk / demodulate_qpsk), imported from /root/reference;
 * remove_prefix_cc / extract_burst_cc: these are GNU Radio blocks (no GNU Radio in this image), so the
   expected values come from a line-by-line NumPy transcription of their general_work loops
   (lib/remove_prefix_cc_impl.cc:84-115, lib/extract_burst_cc_impl.cc:117-242) in float64, plus the
   reference's own QA scenarios (python/qa_extract_burst_cc.py:34-82, python/qa_remove_prefix_cc.py:40-84,
   the latter with pygfdm-generated frames).

usage:  python tests/golden/make_golden_next.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_shims  # noqa: E402


def extract_burst_general_work(stream, burst_len, backoff, starts, scales, rotations, cfo, max_bursts):
    """lib/extract_burst_cc_impl.cc:117-242 with the tags as arrays; float64/complex128 arithmetic,
    CFO rotation exact (inc**i) instead of VOLK's recursive rotator."""
    avail = len(stream)
    noutput = max_bursts * burst_len
    consumed, produced = avail, 0
    out = []
    for t, burst_start in enumerate(starts):
        actual_start = burst_start - backoff
        if avail - burst_start >= burst_len and produced + burst_len <= noutput:
            scale = 1.0 if scales is None else float(scales[t])
            if actual_start < 0:
                z = min(-actual_start, burst_len)
                b = np.concatenate((np.zeros(z, dtype=complex), scale * stream[:burst_len - z]))
            else:
                b = scale * stream[actual_start:actual_start + burst_len]
            if cfo:
                pr = 1.0 + 0.0j if rotations is None else complex(rotations[t])
                inc = np.conj(pr) / abs(pr)
                b = b * inc ** np.arange(burst_len)
            out.append(b)
            produced += burst_len
            consumed = burst_start + burst_len
        else:
            consumed = max(0, burst_start)
            break
    out = np.array(out, dtype=complex).reshape(-1, burst_len)
    return out, consumed


def main():
    install_shims()
    sys.path.insert(0, REF)
    from pygfdm import symbolmapping, utils
    from pygfdm.cyclic_prefix import pinch_cp_add_block
    from pygfdm.gfdm_modulation import modulate_mapped_gfdm_block
    from pygfdm.mapping import get_subcarrier_map
    from pygfdm.preamble import mapped_preamble

    g = {}
    rng = np.random.RandomState(20260817)

    # ---- symbol mapping (pygfdm) --------------------------------------------------------------
    for order in (1, 2):
        const = symbolmapping.generate_constellation(order)
        bits = rng.randint(0, 2, 96 * order)
        syms = symbolmapping.bits2symbols(bits, const).ravel()  # pack_bits keeps a trailing axis of 1
        noisy = syms + 0.3 * (rng.standard_normal(syms.size) + 1j * rng.standard_normal(syms.size))
        g['sm%d_points' % order] = np.asarray(const, dtype=complex)
        g['sm%d_bits' % order] = bits.astype(np.uint8)
        g['sm%d_chunks' % order] = symbolmapping.pack_bits(bits, order).ravel().astype(np.uint8)
        g['sm%d_symbols' % order] = syms
        g['sm%d_noisy' % order] = noisy
        g['sm%d_noisy_bits' % order] = symbolmapping.symbols2bits(noisy, const).astype(np.uint8)
    q = utils.get_random_qpsk(200, seed=4711)
    qn = q + 0.4 * (rng.standard_normal(q.size) + 1j * rng.standard_normal(q.size))
    g['qpsk_syms'] = qn
    g['qpsk_demod_bits'] = utils.demodulate_qpsk(qn).astype(np.uint8)  # [re<0, im<0] per symbol

    # ---- remove_prefix_cc: the scenario of python/qa_remove_prefix_cc.py:40-84 (smaller) ---------
    n_frames, timeslots, subcarriers, active = 6, 9, 32, 26
    cp_len = subcarriers // 2
    smap = get_subcarrier_map(subcarriers, active)
    preamble, _ = mapped_preamble(4711, 'rrc', .5, active, subcarriers, smap, 2, cp_len, cp_len // 2)
    block_len = timeslots * subcarriers
    offset = len(preamble) + cp_len
    frame_len = len(preamble) + block_len + cp_len
    data, ref = [], []
    np.random.seed(99)
    for _ in range(n_frames):
        d_block = modulate_mapped_gfdm_block(utils.get_random_qpsk(timeslots * active), timeslots, subcarriers,
                                             active, 2, .5)
        frame = np.concatenate((preamble, pinch_cp_add_block(d_block, timeslots, subcarriers, cp_len, cp_len // 2)))
        assert len(frame) == frame_len
        ref.append(frame[offset:offset + block_len])
        data.append(frame)
    g['rp_params'] = np.array([frame_len, block_len, offset])
    g['rp_in'] = np.array(data)
    g['rp_out'] = np.array(ref)

    # ---- extract_burst_cc ------------------------------------------------------------------------
    # (a) python/qa_extract_burst_cc.py:34-82 with fewer frames
    n_frames, burst_len, gap_len = 12, 383, 53
    data = np.arange(burst_len, dtype=complex)
    ref, starts = [], []
    for i in range(n_frames):
        frame = np.ones(burst_len) * (i + 1)
        ref.append(frame)
        starts.append(burst_len + i * (burst_len + gap_len))
        data = np.concatenate((data, frame, np.zeros(gap_len)))
    g['eb_qa_params'] = np.array([burst_len, 0])
    g['eb_qa_in'] = data
    g['eb_qa_starts'] = np.array(starts, dtype=np.int64)
    g['eb_qa_out'] = np.array(ref)
    # (b) backoff, negative actual start, scale factors, a tag too close to the end, limited output space
    burst_len, backoff = 200, 24
    stream = rng.standard_normal(3000) + 1j * rng.standard_normal(3000)
    starts = np.array([10, 300, 777, 1500, 2300, 2850], dtype=np.int64)
    scales = np.array([0.5, 1.25, 2.0, 0.75, 1.0, 3.0])
    rots = np.exp(1j * np.array([0.0, 1e-3, -2.5e-3, 0.01, -0.02, 0.005])) * np.array([1.0, 2.0, 0.5, 1.0, 3.0, 1.0])
    g['eb_params'] = np.array([burst_len, backoff])
    g['eb_in'] = stream
    g['eb_starts'] = starts
    g['eb_scales'] = scales
    g['eb_rots'] = rots
    for name, cfo, mb in (('plain', False, 6), ('cfo', True, 6), ('short', False, 3)):
        out, consumed = extract_burst_general_work(stream, burst_len, backoff, starts, scales, rots, cfo, mb)
        g['eb_%s_out' % name] = out
        g['eb_%s_consumed' % name] = np.array([consumed])

    np.savez_compressed(os.path.join(HERE, 'next_rows.npz'), **g)
    print('wrote next_rows.npz:', {k: v.shape for k, v in g.items()})


if __name__ == '__main__':
    main()
