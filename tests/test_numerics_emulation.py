"""The fp32 op-by-op emulation behind DESIGN.md's account of the near-mode gradient error (scripts/ex2_bias_emulation.py)
stays runnable and keeps telling the same story: K1d's formulation with exactly rounded functions is accurate to ~1e-6
of |g|inf near the mode; a 5e-8 mean relative error of the exponential (MUFU.EX2) multiplies that several times."""

import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ex2_bias_emulation_story():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ex2_bias_emulation.py"), "60000"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    err = {m.group(1).strip(): float(m.group(2))
           for m in re.finditer(r"^(.*?)\s+gradient error / \|g\|inf = ([0-9.e+-]+)$", out.stdout, re.M)}
    exact = err["K1d formulation, exact functions"]
    biased = err["K1d formulation, ex2 bias -5e-8 (x2 < 0) / +3e-8 (x2 > 0)"]
    engine = err["engine formulation, exact functions"]
    assert exact < 5e-6 and engine < exact, err
    assert biased > 2.0 * exact, err
