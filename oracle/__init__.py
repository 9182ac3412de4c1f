"""CPU oracle for the GCBF+ hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain torch-CPU / NumPy restatement of the reference's
(MIT-REALM/gcbfplus, JAX) algorithm for the hot path named in BASELINE.json.
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``gcbfplus_b200`` never imports it and fails loudly when
its CUDA library is missing.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures with
inputs/outputs (only pretrained parameter pickles), and JAX/Flax/jraph/optax are
not installable in this image, so the reference itself cannot be executed to pin
this restatement.  What *is* pinned (tests/test_oracle.py): the LQR gains
(SURVEY 8a6), the parameter tree / counts of the 8 pretrained pickles, the
dense (reference-layout) == sparse equivalence, closed-form geometry cases, the
QP labels against SciPy's SLSQP and the KKT conditions (oracle/qp.py), and the
random draws of reset against Random123's Threefry vectors and the values jax
prints for split / uniform of PRNGKey(0) and PRNGKey(42) (oracle/reset.py).

Modules: geometry (obstacles, LiDAR), envs (graphs, dynamics, masks), nn (GNN),
algo (act, losses, optimizer, rollout), qp (action labels), reset (scenarios).

Every function cites the reference file:line it restates
(paths relative to /root/reference/).
"""
