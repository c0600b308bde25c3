"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the iqwaveform spectral-analysis hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU arm).  The product package ``iqwaveform_b200`` never imports it and fails
loudly when its CUDA library is missing.

Contents
--------
``iqw_oracle``   numpy/scipy restatement of the reference algorithm (each function cites the
                 reference file:line it follows).  Pinned bit-for-bit against the real reference
                 imported through ``ref_shim`` (``tests/test_oracle_vs_reference.py`` here, and the
                 committed fixtures under ``tests/golden/`` on the GPU box, where ``/root/reference``
                 does not exist).
``ref_shim``     imports the UNMODIFIED reference from ``/root/reference/src`` (this container
                 only) behind five stub modules; used to generate the golden fixtures.
``make_golden``  the script that wrote ``tests/golden/*.npz``.

Parity pin status: the reference's own tests hold no vector for this path, so the
pin is "outputs of the reference itself run here" -- see ``make_golden.py``.
"""
