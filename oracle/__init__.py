"""CPU oracle for the s2v / GNN-embedding hot path of rvdweerd/simmodel.

TEST INFRASTRUCTURE ONLY.  Nothing under ``simmodel_b200/`` may import this
package.  The only allowed users are ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there
only as the checker / the timed CPU baseline, never as the product path.

Parity status: PINNED.  The reference has no tests or golden vectors for this
path, but its Python modules import in the build container
once the absent third-party packages are stubbed (``oracle/ref_import.py``).
``oracle/gen_golden.py`` runs the *unmodified* reference classes
(``modules/gnn/comb_opt.py:QNet``, ``modules/ppo/models_basic_lstm.py:
PPO_GNN_Single_LSTM``) on seeded inputs, asserts that this restatement agrees
with them, and commits the input/output/gradient vectors under
``tests/golden/``.
"""
