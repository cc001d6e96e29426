"""DESIGN.md section 6 claims that one channel cannot be split in time: a biquad restarted inside a stream (zeroed output state,
correct input history) never rejoins the true trajectory, because the 14-bit residual is an exact carry.  This keeps the
experiment behind that claim (tools/study/biquad_state_merge.c, a plain-C restatement of filter_biquad.cpp:56-63) runnable."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COEFS = "236552419 473104839 236552419 175469220 -47937074 1049016272 -1483533003 1049016272 1483533003 -1024290721".split()


def test_restarted_biquad_never_rejoins(tmp_path):
    exe = str(tmp_path / "merge")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "study", "biquad_state_merge.c"), "-lm"], check=True)
    for kind, trials in ((2, 6), (0, 4)):  # tone + noise, noise
        out = subprocess.run([exe] + COEFS + [str(kind), "3000", str(trials)], capture_output=True, text=True, check=True, timeout=300).stdout
        m = re.search(r"stage1 never (\d+) .*cascade never (\d+)", out)
        assert m, out
        assert int(m.group(1)) == 4 * trials and int(m.group(2)) == 4 * trials, out
    # silence is the one input for which both trajectories coincide from the start (nothing to carry)
    out = subprocess.run([exe] + COEFS + ["3", "0", "3"], capture_output=True, text=True, check=True, timeout=300).stdout
    assert "cascade never 0" in out, out
