"""CPU oracle for the TSADAR form-factor hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain NumPy float64 (values) and torch float64 (autograd
gradients), the algorithm of the reference path

    tsadar/core/physics/form_factor.py, ratintn.py, generate_spectra.py, irf.py,
    tsadar/core/thomson_diagnostic.py, tsadar/inverse/loss_function.py (loss stage)

following the reference line by line (each function cites the file:line it follows).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU baseline.
The product (``tsadar_b200``) never imports it and has no CPU fallback.

Parity pin: ``tests/golden/ThryE-1d.npy`` (the reference's own golden vector
``tests/test_forward/ThryE-1d.npy``) is reproduced by ``oracle.np_oracle`` through the full
diagnostic (see ``tests/test_oracle_golden.py``).  The 2V (ARTS-2V) path and the ATS IRF have
no surviving golden in the reference checkout: for those rows the oracle is "parity unpinned".
"""
