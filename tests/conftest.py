import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def report(name, **values):
    """One line per parity margin: printed (pytest -s / -rP) and appended to gpurun_out/parity_margins.txt, which the
    builder copies to profiles/ so that the margins of the shipped build are on record."""
    def fmt(v):
        if isinstance(v, (bool, np.bool_)):
            return "yes" if v else "no"
        if isinstance(v, (float, np.floating)):
            return f"{float(v):.4g}"
        return str(v)
    line = f"PARITY {name}: " + ", ".join(f"{k}={fmt(v)}" for k, v in values.items())
    print(line)
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_margins.txt"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def unpack_notes(z, prefix):
    return [{"pitch": int(p), "onset": float(a), "offset": float(b), "velocity": int(v)}
            for p, a, b, v in zip(z[prefix + "_pitch"], z[prefix + "_onset"], z[prefix + "_offset"],
                                  z[prefix + "_velocity"])]


NOTE_CASES = ["smooth", "offset_saturated", "plateaus", "white", "all_zero", "all_one", "one_frame", "two_frames",
              "velocity_zero", "dense_sigmoid"]


def note_variants(z, name):
    """(mode_offset, mode_velocity, golden notes) for every variant of a case stored in notes.npz."""
    out = []
    for mo in ("shorter", "longer", "offset"):
        for mv in ("ignore_zero", "org"):
            key = f"{name}__{mo}__{mv}"
            if key + "_pitch" in z.files:
                out.append((mo, mv, unpack_notes(z, key)))
    return out
