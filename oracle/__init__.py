"""CPU oracle for the MusicStyleTransfer hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy / torch-CPU fp32, plus a plain-C
rasteriser in ``raster.c``) of the reference algorithms on the hot path
(SURVEY.md §8(a), rows A1–A12).  Every function cites the reference file:line
it follows (paths relative to ``/root/reference/music_style_transfer``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import or execute anything under ``oracle/``.
The product package ``musicstyletransfer_b200`` never does; it fails loudly when
its CUDA library is missing.

Parity pinning status:
  * A1 (event tokens)            PINNED  — golden vectors in tests/golden/ were
    produced by executing the reference's own ``EventBasedMIDIReader._parse_track``
    (tests/golden/make_golden.py, with a stub ``midi`` module) on the 37 fixtures.
  * A2 (row chunking)            PINNED  — reference ``MelodyDataset._get_token_arrays``
    executed over a NumPy-backed ``mxnet`` shim (same script).
  * A8/A9/A10 (loss formulas)    PINNED formulas — reference ``loss.py`` executed over
    the same shim.
  * A3–A7 forward                PINNED to the reference source executed over the shim
    (shim = our statement of MXNet-1.3 operator semantics; MXNet itself cannot run here).
  * A11 (backward, Adam), A12 (RNG), piano roll — "parity unpinned": no reference
    artefact exists at that boundary; the oracle is the specification.
"""
