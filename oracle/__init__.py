"""CPU oracle for the vit-gan hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU and in plain ``torch`` fp32/fp64 ops, the
algorithm of the reference's ViT generator / discriminator blocks and of the
G+D training step (reference: ``src/v2/modules.py``, ``src/v2/training.py``,
``src/v1/*.py``; every function cites the file:line it follows).

It exists only so that the CUDA path can be *checked*:

* ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
  ``--impl reference`` legs of ``bench.py`` are the only allowed importers.
* Nothing under ``vit-gan_b200/`` (the product) imports it; the product fails
  loudly when the CUDA extension is missing instead of falling back to this.

Parity pinning: the reference ships **no** tests, golden vectors or fixtures
for this path ("parity unpinned" by the reference itself).
The oracle is therefore pinned against outputs of the reference *itself*,
executed in the build container by ``oracle/make_golden.py`` (imports the real
``/root/reference`` modules, commits small fixtures under ``tests/golden/``)
and, whenever ``/root/reference`` is present, checked live and bit-exactly by
``tests/test_oracle_vs_reference.py``.
"""

from . import v2, v1, harness  # noqa: F401
