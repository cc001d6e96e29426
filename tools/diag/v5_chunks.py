"""Diagnostic: msdr_chain_update (host buffers, channel chunks of 64) on the row-block kernel: where do mismatches sit?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import minimal_sdr_b200 as m
import oracle_lib as ol
from chain_helpers import configure_pair

K = m.load_ref_constants()
orc = ol.CheckerLib("orc")
modes = m.synth.mixed_modes(203)
x = m.synth.batch(modes, 128 * 9)
for variant, chunk in ((4096, 64), (4096, 0), (4096 + 1, 64), (0, 64)):
    g, o = configure_pair(m, orc, K, modes)
    g.set_option("variant", variant)
    if chunk:
        g.set_option("host_chunk_channels", chunk)
    y = g.update(x)
    want = o.run(x)[0]
    bad = np.argwhere(y != want)
    print(f"variant {variant} chunk {chunk}: {len(bad)} mismatches, kernel {g.last_kernel()[:24]}")
    if len(bad):
        chs = sorted(set(bad[:, 0]))
        print("  channels", chs[:16], "...", len(chs))
        for c in chs[:3]:
            idx = bad[bad[:, 0] == c][:, 1]
            print(f"  ch {c} mode {modes[c]}: first {idx[:8]}, count {len(idx)}, by index mod 64 >= 48: {(idx % 64 >= 48).sum()}, max |diff| {np.abs(y[c].astype(int) - want[c]).max()}")
    g.close(); o.close()
