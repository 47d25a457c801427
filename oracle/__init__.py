"""CPU oracle for the BoT GATConv hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``bot_b200/`` may import this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and only as the checker / the timed CPU arm.

PARITY UNPINNED: the reference (AiRyunn/BoT) has no tests or golden vectors and
its sparse arithmetic lives in DGL 0.5.* (README.md:9), which is neither vendored
under /root/reference nor installable offline.  The oracle restates the math of
``src/no-sampling/models.py:475-566`` and ``src/ogbn-proteins/models.py:87-168``
with DGL 0.5 op semantics and is self-validated by fp64
gradcheck, hand-computed cases and invariants (tests/test_oracle.py).
"""
