"""CPU oracle for the tfep MAF / (T)FEP hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, with plain PyTorch-CPU / numpy operations, the algorithm of the
reference (andrrizzi/tfep, mounted read-only at /root/reference in the build
container) for the path this repository accelerates:

    tfep.nn.flows.MAF forward / inverse with log|det J|   (flow_oracle.py)
    tfep.analysis.fep_estimator and tfep.analysis.bootstrap (analysis_oracle.py)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never from ``tfep_b200`` itself.

Parity pinning: the restatement is checked bit-for-bit (fp32 and fp64) against
the real reference, imported in the build container by ``ref_import.py`` (see
``check_against_reference.py``), and against the golden vectors that
``make_golden.py`` generated from the real reference and committed under
``tests/golden/``.  The reference holds no stored golden vectors of its own for
this path; its known-answer tables
(tfep/tests/nn/conditioners/test_made.py:31-67) are replayed in
``tests/test_oracle_golden.py``.
"""
