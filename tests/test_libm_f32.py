"""csrc/libm_f32.h (float-only atanf / atan2f / acosf used by the projection kernels) equals the C
library the reference's numba code calls (visibility.py:167-168), bit for bit.  Compiled for the
host with gcc -ffp-contract=off; the device build routes the same source through the
round-to-nearest intrinsics.  The exhaustive run (stride 1: 6.4e9 values, 0 mismatches) is recorded
in the header; this test samples every 997th bit pattern plus 4e6 atan2f pairs."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_libm_f32_matches_c_library(tmp_path):
    exe = tmp_path / "libm_f32_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe),
                    os.path.join(ROOT, "tests", "native", "libm_f32_check.c"), "-lm"], check=True)
    out = subprocess.run([str(exe), "997", "4000000"], check=True, capture_output=True, text=True).stdout
    rows = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out.splitlines()}
    assert set(rows) == {"acosf", "atanf", "atan2f"}
    for name, (bad, tot) in rows.items():
        assert tot > 1_000_000 and bad == 0, (name, bad, tot)
