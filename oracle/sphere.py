"""numpy restatement of the sphere-geometry helpers of Sphere_Grad_Descent.py (alias ``SGD``).  TEST INFRASTRUCTURE ONLY.

These are rows C1-C3 of SURVEY.md section 8(a); the CUDA path fuses each into
single-pass reduction + axpy kernels behind the DevVec type.  Parity unpinned - see oracle/__init__.py.
"""
import numpy as np


def transport_vector(X_k, dkm1, inner_prod, args_IP=(), kwargs_IP={}):
    """SGD:625-642:  d - (<X,d>/<X,X>) X"""
    L2 = np.sqrt(inner_prod(X_k, X_k, *args_IP, **kwargs_IP))
    return dkm1 - (inner_prod(X_k, dkm1, *args_IP, **kwargs_IP) / (L2 ** 2)) * X_k


def tangent_vector(X_k, Nab_Jk, inner_prod, args_IP=(), kwargs_IP={}):
    """SGD:644-659:  grad - (<X,grad>/<X,X>) X"""
    return Nab_Jk - (inner_prod(X_k, Nab_Jk, *args_IP, **kwargs_IP) / inner_prod(X_k, X_k, *args_IP, **kwargs_IP)) * X_k


def Update_vector(X_k, alpha_k, d_k, M_0, inner_prod, args_IP=(), kwargs_IP={}):
    """SGD:661-690:  retraction  f = X + alpha d;  f * sqrt(M_0/<f,f>)"""
    f = X_k + alpha_k * d_k
    L2_f = inner_prod(f, f, *args_IP, **kwargs_IP)
    return f * np.sqrt(M_0 / L2_f)
