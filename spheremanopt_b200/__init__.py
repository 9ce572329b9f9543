"""spheremanopt_b200 - B200-native (sm_100a, fp64 CUDA) hot path of mannixp/SphereManOpt.

The f / Grad_f / Inner_Product callables of the two periodic-Fourier examples, behind the reference's own
function names and argument lists, so that the unmodified ``Optimise_On_Multi_Sphere`` / ``Adjoint_Gradient_Test``
can call them:

    from spheremanopt_b200 import sh23   # mirrors Swift_Hohenberg/FWD_Solve_SH23.py
    from spheremanopt_b200 import kdyn   # mirrors Kinematic_Dynamo/FWD_Solve_KDyn.py

All arithmetic runs in libsmo_b200.so (include/smo_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
