"""CPU oracle for the GCBF+ hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain torch-CPU / NumPy restatement of the reference's
(MIT-REALM/gcbfplus, JAX) algorithm for the hot path named in BASELINE.json.
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``gcbfplus_b200`` never imports it and fails loudly when
its CUDA library is missing.

PARITY UNPINNED for the network / loss outputs: the reference ships no tests, golden vectors or fixtures with
inputs/outputs (only pretrained parameter pickles), and JAX/Flax/jraph/optax are not installable in this image
(re-probed on every run: tests/helpers.py probe_reference_stack), so the reference itself cannot be executed here
to pin this restatement.  Staged for the day it can: tests/golden/make_io_fixtures.py (runs the UNMODIFIED
reference on BASELINE configs 1-3 -> tests/golden/ref_io_*.npz) and tests/test_reference_goldens.py (checks this
oracle AND the CUDA path against those files; skips loudly while they are absent).

What *is* pinned today (tests/test_oracle.py): the LQR gains, the parameter tree / counts of the 8
pretrained pickles, the dense (reference-layout) == sparse equivalence, closed-form geometry cases, the QP labels
against SciPy's SLSQP and the KKT conditions (oracle/qp.py), and the random draws of reset in BOTH threefry stream
layouts (jax_threefry_partitionable off / on) against Random123's Threefry vectors and the values JAX's own
documentation prints: split / uniform / normal of PRNGKey(0) / key(42) and the tutorial's split-and-draw loop
(oracle/reset.py).

Modules: geometry (obstacles, LiDAR), envs (graphs, dynamics, masks), nn (GNN),
algo (act, losses, optimizer, rollout), qp (action labels), reset (scenarios).

Every function cites the reference file:line it restates
(paths relative to /root/reference/).
"""
