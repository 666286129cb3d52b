"""The note stage's chunked algorithm on the CPU (no GPU needed): etude_b200/csrc/notes.cuh's per-item functions are
__host__ __device__; tests/notes_host.cu compiles them for the host and restates the kernels' orchestration serially.
Checked bit-exactly against the reference-generated goldens and against the oracle's C restatement on random rolls with
plateaus, saturation, sparse onsets and ragged lengths -- the cases where a note's neighbours lie in other chunks."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import NOTE_CASES, ROOT, note_variants

SRC = os.path.join(ROOT, "tests", "notes_host.cu")
LIB = os.path.join(ROOT, "tests", "libnotes_host.so")
DT = np.dtype([("pitch", np.int32), ("velocity", np.int32), ("onset", np.float64), ("offset", np.float64)])
MV = {"ignore_zero": 0, "org": 1}
MO = {"shorter": 0, "longer": 1, "offset": 2}


@pytest.fixture(scope="module")
def host():
    deps = [SRC, os.path.join(ROOT, "etude_b200", "csrc", "notes.cuh"), os.path.join(ROOT, "etude_b200", "csrc", "common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-shared", "-Xcompiler",
                        "-fPIC,-ffp-contract=off", "-o", LIB, SRC], check=True)
    lib = ctypes.CDLL(LIB)
    lib.notes_host.restype = ctypes.c_int64
    lib.notes_host.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                                        ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def run(on, off, mpe, vel, t_on, t_off, t_mpe, mode_velocity="ignore_zero", mode_offset="shorter"):
        on, off, mpe = (np.ascontiguousarray(a, np.float32) for a in (on, off, mpe))
        vel = np.ascontiguousarray(vel, np.int8)
        T = on.shape[0]
        out = np.zeros(T * 88, DT)
        n = lib.notes_host(on.ctypes.data, off.ctypes.data, mpe.ctypes.data, vel.ctypes.data, T, 21, 256 / 16000, t_on, t_off, t_mpe,
                           MV[mode_velocity], MO[mode_offset], out.ctypes.data)
        r = out[:n]
        return [{"pitch": int(p), "onset": float(a), "offset": float(b), "velocity": int(v)}
                for p, a, b, v in zip(r["pitch"].tolist(), r["onset"].tolist(), r["offset"].tolist(), r["velocity"].tolist())]
    return run


@pytest.mark.parametrize("case", NOTE_CASES)
def test_chunked_notes_equal_reference_goldens(host, golden, case):
    z = golden("notes")
    t_on, t_off, t_mpe = z[case + "__thr"]
    for mo, mv, ref in note_variants(z, case):
        got = host(z[case + "__onset"], z[case + "__offset"], z[case + "__mpe"], z[case + "__velocity"], t_on, t_off, t_mpe, mv, mo)
        assert got == ref, (case, mo, mv)


def test_chunked_notes_equal_oracle_on_random_rolls(host):
    """Lengths around the chunk size, plateaus (equal values), saturated offsets, sparse and dense onsets, velocity zeros,
    every mode: the chunked walk must equal the serial oracle bit for bit."""
    from oracle import notes as onotes
    rng = np.random.default_rng(11)
    for t in (1, 2, 255, 256, 257, 511, 512, 513, 1000, 2048 + 17):
        for density in (0.02, 0.5, 1.0):
            q = int(rng.integers(2, 30))
            on = (np.round(rng.random((t, 88)) * q) / q).astype(np.float32)
            on[rng.random((t, 88)) > density] = 0.0                               # sparse onsets: neighbours far away
            off = np.minimum(1.0, rng.random((t, 88)) * (1.0 + 0.5 * density)).astype(np.float32)
            if density < 0.1:
                off[:] = np.float32(0.3)                                          # no offset peak anywhere
            mpe = rng.random((t, 88)).astype(np.float32)
            mpe[:, ::3] = 0.9                                                     # pitches whose notes never end by mpe
            vel = rng.integers(0, 4, (t, 88)).astype(np.int8)                     # many zeros
            for mv in ("ignore_zero", "org"):
                for mo in ("shorter", "longer", "offset"):
                    got = host(on, off, mpe, vel, 0.5, 1.0, 0.5, mv, mo)
                    ref = onotes.mpe2note(on, off, mpe, vel, 0.5, 1.0, 0.5, mode_velocity=mv, mode_offset=mo)
                    assert got == ref, (t, density, mv, mo, len(got), len(ref))


def test_chunked_notes_long_plateaus_across_chunks(host):
    """Plateaus longer than a chunk (all frames of a qualifying run are peaks) and saturated rolls."""
    from oracle import notes as onotes
    t = 1500
    on = np.zeros((t, 88), np.float32)
    on[100:900, 0] = 0.8                      # one 800-frame plateau over four chunks
    on[250:260, 1] = 0.7
    on[255:257, 2] = 0.9                      # two-frame plateau straddling the first chunk border (ulp-inverted peak times)
    on[:, 3] = 1.0                            # saturated everywhere
    off = np.zeros((t, 88), np.float32)
    off[300:1200, 0] = 1.0
    off[:, 3] = 1.0
    mpe = np.full((t, 88), 0.6, np.float32)
    mpe[700:, 0] = 0.1
    vel = np.full((t, 88), 64, np.int8)
    vel[255, 2] = 0
    for mo in ("shorter", "longer", "offset"):
        got = host(on, off, mpe, vel, 0.5, 1.0, 0.5, "ignore_zero", mo)
        assert got == onotes.mpe2note(on, off, mpe, vel, 0.5, 1.0, 0.5, mode_offset=mo), mo
