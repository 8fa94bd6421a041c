"""upright_b200 — batched, B200-native MPC solve for the waiter's problem.

Module map (reference name -> here):
    upright_core.parsing            -> upright_b200.config, upright_b200.objects
    upright_core.math / polyhedron  -> upright_b200.geometry
    upright_control.wrappers        -> upright_b200.settings
    upright_control.manager         -> upright_b200.manager
    upright_control.trajectory      -> upright_b200.trajectory
    upright_control.bindings        -> upright_b200.bindings + csrc/ (C ABI, CUDA)
"""
__version__ = "0.1.0"
