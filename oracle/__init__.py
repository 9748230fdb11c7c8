"""CPU oracle for the PINN residual hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``torchphysics_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and there only as the checker / the timed CPU arm.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the real reference
(boschresearch/torchphysics v1.1.0, imported from ``/root/reference/src`` with the two
import stubs in ``oracle/stubs``) and stores its outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those files.

Where the arithmetic lives: the reference has no numerical kernels of its own -- all
arithmetic happens in PyTorch (``torch>=2.0.0``, reference ``pyproject.toml:23``; 2.11.0
here): ``nn.Linear``/``nn.Tanh`` (reference ``models/fcn.py:9-29``) and
``torch.autograd.grad(create_graph=True)`` (reference
``utils/differentialoperators/differentialoperators.py:34,40,64,160,254``).
``ref_autograd`` restates the reference's *call sites* on top of the same torch CPU ops,
``jet_numpy`` restates the mathematics independently in float64 numpy (forward-Laplacian
propagation + hand-written reverse sweep), ``geometry`` restates domain membership,
volumes and deterministic grids in numpy.
"""
