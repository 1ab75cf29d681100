"""the strict kernels' transcendental functions (vkdt_b200/csrc/kernels/libm_exact.h) against this machine's libm, on the
host: the header compiled for the host with -ffp-contract=off (tests/tools/libm_exact_check.c).  every 4099th float bit pattern for
expf / exp2f / logf / log2f and 2 x 2^21 random + 19 x 2^26 structured pairs for powf here; the exhaustive run (all 2^32
arguments, 2 x 10^9 pairs: 0 mismatches, 12 core-minutes) is recorded in DESIGN.md section 4."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libm_exact_header_matches_libm(tmp_path):
    exe = str(tmp_path / "lmcheck")
    subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", os.path.join(ROOT, "tests", "tools", "libm_exact_check.c"),
                    "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "4099", str(1 << 21)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    counts = dict(l.split() for l in r.stdout.splitlines())
    assert counts == {"expf": "0", "exp2f": "0", "logf": "0", "log2f": "0", "powf": "0"}, counts
