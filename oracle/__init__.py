"""TEST INFRASTRUCTURE ONLY — CPU oracle of the reference's tiled-prediction path.

Nothing under ``oracle/`` is imported by the product package ``bio_image_unet_b200``. Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Parity status: PINNED. The reference's own tests hold no golden vectors (utils/test.py:18-111 are assertion-free
smoke runs on unseeded data), so the oracle is pinned against outputs of the *unmodified reference itself*,
executed in the authoring container by ``tests/golden/make_golden.py`` (import recipe: ``oracle/ref_import.py``)
and committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
