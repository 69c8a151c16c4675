"""CPU oracle for the pyNeuralEMPC NLP-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyneuralempc_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and there only as the
checker or as the timed CPU baseline -- never as the thing shipped.

Parity status: the reference ships no tests, golden vectors or known-answer
fixtures for this path, and its derivative arithmetic
lives in TensorFlow / JAX, which are not installed here.  The oracle is therefore
pinned as follows (see DESIGN.md "Oracle"):

* ``dense_ref`` / ``blocks_np`` (integrators, IPOPT callback glue) are checked
  against the *unmodified reference code* imported from ``/root/reference``
  through ``oracle.shim`` -- outputs committed as ``tests/golden/*.npz`` by
  ``tests/golden/make_golden.py``.
* ``mlp_np`` (the MLP forward / Jacobian / per-output Hessian that TensorFlow
  autodiff produces in the reference, ``model/tensorflow.py:49-109``) is
  "parity unpinned" against TensorFlow itself (tensorflow is absent, unpinned in
  ``setup.py:20``); it is pinned instead against ``torch.func`` autodiff and
  central finite differences in ``tests/test_oracle_mlp.py``.

Modules
-------
mlp_np         analytic tanh-MLP forward / Jacobian / per-output Hessian and the
               dense reference layouts of ``KerasTFModel``.
dense_ref      literal dense restatement of Discret/Unity/RK4 integrators, the
               numeric structure probing and the ``IpoptProblem`` callbacks.
blocks_np      per-step block formulation (O(H) memory) + sparse assembly.
objectives_np  the three cost families used with ``JAXObjectifFunc``.
structure      analytic Jacobian / Hessian sparsity in reference ordering.
shim           imports the real reference package with tensorflow/jax/cyipopt
               stubbed (only works where ``/root/reference`` exists).
"""
