"""Stub so the UNMODIFIED reference Sphere_Grad_Descent.py imports without MPI (it only reads
MPI.COMM_WORLD.rank, SGD:3, SGD:822).  Test infrastructure only."""


class _Comm:
    rank = 0
    size = 1

    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0


class MPI:
    COMM_WORLD = _Comm()
