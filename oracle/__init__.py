"""CPU oracle for the WordGesture-GAN training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
CPU arm being timed.  The product (``wordgesture-gan_b200``) never imports it
and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no golden vectors, so
the oracle is pinned against outputs of the reference itself, generated in the
build container by ``oracle/make_golden.py`` (which imports the unmodified
reference from /root/reference) and committed under ``tests/golden/``.  The
generating script asserts oracle == reference (fp64) before writing a fixture.
"""
