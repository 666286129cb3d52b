"""Oracle: piano-rolls -> notes (TEST INFRASTRUCTURE, see oracle/__init__.py).

ctypes wrapper over ``mpe2note.c`` (restating etude/data/extractor.py:256-418)
plus a restatement of ``_note2json`` (extractor.py:432-446).
"""
import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HOP_SEC = float(256 / 16000)
NOTE_MIN = 21


class _Note(ctypes.Structure):
    _fields_ = [("pitch", ctypes.c_int32), ("velocity", ctypes.c_int32),
                ("onset", ctypes.c_double), ("offset", ctypes.c_double)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_notes.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.oracle_mpe2note.restype = ctypes.c_int64
        lib.oracle_mpe2note.argtypes = [ctypes.c_void_p] * 4 + [
            ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.POINTER(_Note))]
        lib.oracle_free.argtypes = [ctypes.c_void_p]
        _LIB = lib
    return _LIB


_MODE_VELOCITY = {"ignore_zero": 0, "org": 1}
_MODE_OFFSET = {"shorter": 0, "longer": 1, "offset": 2}


def mpe2note(a_onset, a_offset, a_mpe, a_velocity, thred_onset=0.5, thred_offset=0.5, thred_mpe=0.5,
             mode_velocity="ignore_zero", mode_offset="shorter", hop_sec=HOP_SEC, note_min=NOTE_MIN):
    """Same signature/defaults as AMTAPC_Extractor._mpe2note (extractor.py:256)."""
    on = np.ascontiguousarray(a_onset, dtype=np.float32)
    off = np.ascontiguousarray(a_offset, dtype=np.float32)
    mpe = np.ascontiguousarray(a_mpe, dtype=np.float32)
    vel = np.ascontiguousarray(a_velocity, dtype=np.int8)
    t, num_note = on.shape
    out = ctypes.POINTER(_Note)()
    n = _lib().oracle_mpe2note(on.ctypes.data, off.ctypes.data, mpe.ctypes.data, vel.ctypes.data, t, num_note,
                               note_min, hop_sec, thred_onset, thred_offset, thred_mpe,
                               _MODE_VELOCITY.get(mode_velocity, 1), _MODE_OFFSET.get(mode_offset, 0),
                               ctypes.byref(out))
    notes = [{"pitch": int(out[i].pitch), "onset": float(out[i].onset), "offset": float(out[i].offset),
              "velocity": int(out[i].velocity)} for i in range(n)]
    _lib().oracle_free(out)
    return notes


def note2json_obj(notes, min_length=0.0):
    """_note2json's filter and key order (extractor.py:432-443)."""
    return [{"onset": n["onset"], "offset": n["offset"], "pitch": n["pitch"], "velocity": n["velocity"]}
            for n in notes if not (n["offset"] - n["onset"] < min_length)]


def note2json(notes, path_output, min_length=0.0):
    with open(path_output, "w", encoding="utf-8") as f:
        json.dump(note2json_obj(notes, min_length), f, ensure_ascii=False, indent=2)
