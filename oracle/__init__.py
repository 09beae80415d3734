"""oracle/ - CPU restatement of the reference (agi-brain/xuance v1.4.4 @ 4f0b05b) rollout->update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``xuance_b200/`` (the product) imports this package.  The only
legal importers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (its ``cpu_baseline`` leg and
its ``--impl reference`` arm) - and there only as the checker / the CPU baseline, never as the thing shipped.

Every function cites the reference file:line it restates.  The restatement is pinned two ways:
  * ``tests/golden/*.npz`` - outputs of the UNMODIFIED reference, generated in the build container by
    ``tests/golden/make_golden.py`` (imports /root/reference with the two import stubs in
    ``oracle/ref_stubs``); the CPU tests check oracle == golden.
  * when /root/reference is present (build container only) ``tests/test_oracle_vs_reference.py`` runs the
    live reference next to the oracle on fresh seeded inputs.
The reference's own tests hold no golden vectors for this path beyond two docstring
known-answers, which are checked too.  Oracle environment of record: torch 2.11.0 (CPU), numpy 2.3.5.
"""
