"""CPU oracle for the Etude Extract-stage hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never on the shipped GPU path.  The product package
``etude_b200`` does not import this package (tests/test_boundary.py checks).

Contents (each function cites the reference file:line it restates):

* ``oracle.logmel``   -- numpy restatement of torchaudio's MelSpectrogram as
  configured by ``etude/data/extractor.py:178-197`` (third-party arithmetic:
  torchaudio 2.6.0 pinned by the reference's requirements.txt; 2.11.0 here).
* ``oracle.model``    -- torch-fp32 functional restatement of
  ``etude/models/amt_apc.py:23-392`` and of the window loop
  ``etude/data/extractor.py:199-253``.
* ``oracle.notes``    -- C restatement (``mpe2note.c``) of
  ``etude/data/extractor.py:256-418`` with this container's NumPy-2 (NEP 50)
  float32/float64 mix, plus ``_note2json`` (432-446).

Parity pinning: the reference holds NO tests and NO golden vectors for this
path.  The oracle is therefore pinned against outputs of
the reference itself, imported from ``/root/reference`` in the build container
by ``oracle/gen_golden.py`` (script committed; fixtures in ``tests/golden/``).
"""
