"""CPU oracle for the temporal neighbour-aggregation path of DyGLib.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dyglib_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  It is a numpy / torch-CPU restatement of the
reference's algorithms (each function cites the reference file:line it follows).

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tests/test_oracle_vs_reference.py`` runs both side by side whenever
``/root/reference`` is importable (it is, in the build container), and
``tests/golden/*.npz`` holds vectors produced by the unmodified reference with
``scripts/make_golden.py`` so the same check travels to the GPU box.
"""
