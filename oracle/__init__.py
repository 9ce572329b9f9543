"""CPU oracle for the SphereManOpt periodic-Fourier hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy fp64 restatement of what the reference's
``f / Grad_f / Inner_Product`` callables compute for the two periodic-Fourier
examples:

* ``oracle.fourier``  - Dedalus-v2 mode layout and transforms (facts [D2-1..5])
* ``oracle.sh23``     - restates ``Swift_Hohenberg/FWD_Solve_SH23.py``
* ``oracle.kdyn``     - restates ``Kinematic_Dynamo/FWD_Solve_KDyn.py``
* ``oracle.sphere``   - restates ``Sphere_Grad_Descent.py:625-690``

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline - never as part of the product path.  The product
(``spheremanopt_b200``) never imports this package and fails loudly when its
CUDA library is missing.

PARITY UNPINNED.  The arithmetic of the reference lives in a third-party,
un-vendored, un-pinned dependency (Dedalus v2 on FFTW/MPI; ``import
dedalus.public as de`` at FWD_Solve_SH23.py:188, FWD_Solve_KDyn.py:198) that is
absent from /root/reference, from this image and from /opt/wheelhouse, and the
reference ships no golden vectors, known-answer tests or fixtures for this path
(SURVEY.md section 8(c)).  The oracle therefore restates Dedalus v2's published
algorithm (SBDF1/CNAB1 multistep IMEX on Fourier pencils, 3/2- and 2-padded
pseudo-spectral products) anchored on the reference's own call sites, and is
pinned only by (i) the reference's own acceptance test - the *unmodified*
``TestGrad.Adjoint_Gradient_Test`` imported from /root/reference gives a
second-order Taylor-remainder slope of 2 (tests/golden/make_golden.py records
the slopes) - (ii) closed-form single-mode linear decay of both time steppers
and (iii) the adjoint dot-product identity.  If a Dedalus install ever becomes
available, re-verify facts [D2-1..13] listed in SURVEY.md section 8(c).
"""
