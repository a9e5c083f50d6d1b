"""CPU oracle for the GAMD hot path (neighbor search -> MDNet forces -> velocity-Verlet).

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this package, and
only as the checker or as the timed CPU baseline.  ``gamd_b200`` never imports it: the
product path fails loudly when the CUDA library is missing instead of falling back here.

What pins this oracle (see DESIGN.md "Oracle and parity pins"):

* ``oracle.model`` (MDNet forward) is checked bit-for-bit (max abs diff 0.0 on torch CPU)
  against the UNMODIFIED reference ``code/nn_module.py`` executed under the pure-torch DGL
  stand-in ``tests/golden/_dgl_stub.py``; the resulting outputs are committed as
  ``tests/golden/*.npz`` by ``tests/golden/make_golden.py``.  -> parity PINNED.
* ``oracle.neighbor.get_neighbor`` (brute-force ``<=`` path) is checked against the
  reference ``code/md_module.py:get_neighbor`` executed here the same way. -> PINNED.
* ``oracle.neighbor.edges_jaxmd`` restates ``code/graph_utils.py:29-61`` (jax-md, which is
  not vendored and not installable here; jax-md version is unpinned upstream).  The
  predicate arithmetic is the published ``space.periodic`` formula evaluated in fp32 with
  one rounding per operation.  -> parity UNPINNED at that third-party boundary; bit-exact
  claims for the cell-list path are against this oracle.
* ``oracle.integrator`` restates the OpenMM ``CustomIntegrator`` programs of
  ``code/hack_integrator.py`` (OpenMM/openmmtools are not installed). -> parity UNPINNED;
  validated by analytic properties (time reversibility, harmonic-oscillator energy).
"""
