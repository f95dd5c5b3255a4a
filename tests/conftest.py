import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = [
    "occu_default", "occu_missing", "occu_5x3", "occu_fp_const", "occu_fp_unocc",
    "rn_default", "rn_5x3", "cop_default", "cop_missing_5x3", "cop_both_fp", "nmix_default", "nmix_missing_5x3",
    "cs_default", "cs_missing_5x3",
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["model"] = str(g["model"])
    g["model_kwargs"] = ast.literal_eval(str(g["model_kwargs"]))
    g["sim_kwargs"] = ast.literal_eval(str(g["sim_kwargs"]))
    g["data"] = dict(site_covs=g["site_covs"], obs_covs=g["obs_covs"], obs=g["obs"])
    if "session_duration" in g:
        g["data"]["session_duration"] = g["session_duration"]
    return g


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return load_golden(request.param)
