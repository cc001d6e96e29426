"""Shared helpers for chain-level parity tests: configure the CUDA chain (through the C ABI) and a checker chain
(oracle or compiled reference) identically, run both, compare bit for bit."""
import numpy as np

import oracle_lib as ol


def tables_for(K, mode):
    if mode in (ol.MODE_USB, ol.MODE_LSB):
        return np.array(K["FIR_SSB_I_coeffs"], np.int16), np.array(K["FIR_SSB_Q_coeffs"], np.int16)
    if mode == ol.MODE_CW:
        return np.array(K["FIR_CW_I_coeffs"], np.int16), np.array(K["FIR_CW_Q_coeffs"], np.int16)
    am = np.array(K["FIR_AM_coeffs_bw2800_fs24000"], np.int16)
    return am, am


def configure_pair(msdr, checker, K, modes, am_q31=False, max_taps=0, biquads="live", tables=None):
    """Returns (gpu_chain, checker_chain) configured alike.  tables: optional {channel: (cI, cQ)} overrides."""
    C = len(modes)
    g = msdr.ReceiveChain(C, max_taps=max_taps, am_q31=am_q31)
    o = checker.chain(C, am_q31)
    c = 0
    while c < C:  # runs of equal mode -> one ranged call each; explicit tables -> per-channel calls
        e = c + 1
        if not tables:
            while e < C and modes[e] == modes[c]:
                e += 1
        cI, cQ = tables[c] if tables and c in tables else tables_for(K, modes[c])
        g.set_mode(modes[c], c, e - c)
        o.set_mode(c, e - c, modes[c])
        g.fir_init(cI, cQ, c, e - c)
        assert o.fir_init(c, e - c, cI, cQ) == 0
        c = e
    if biquads == "live":
        for obj, key in ((0, "biquad1_lowpass_coef"), (1, "biquad2_notch_coef")):
            g.biquad_set_coefficients(obj, 0, K[key])
            o.biquad_set_coefficients(obj, 0, C, 0, K[key])
    elif biquads is not None:
        for obj, stage, coef, ch0, nch in biquads:
            g.biquad_set_coefficients(obj, stage, coef, ch0, nch)
            o.biquad_set_coefficients(obj, ch0, C - ch0 if nch is None else nch, stage, coef)
    return g, o


def run_pair(g, o, x, splits=None):
    """Feed x ([C, n_blocks*128]) to both, optionally in several updates (splits = list of block counts)."""
    nb = x.shape[1] // 128
    splits = splits or [nb]
    assert sum(splits) == nb
    yg, yo, b0 = [], [], 0
    for s in splits:
        part = np.ascontiguousarray(x[:, b0 * 128:(b0 + s) * 128])
        yg.append(g.update(part))
        yo.append(o.run(part)[0])
        b0 += s
    return np.concatenate(yg, axis=1), np.concatenate(yo, axis=1)


def assert_same(a, b, what=""):
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        r, c = bad[0]
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at channel {r} sample {c}: gpu {a[r, c]} vs oracle {b[r, c]}; "
                             f"channels affected {sorted(set(bad[:, 0]))[:10]}")
