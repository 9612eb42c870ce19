"""Import-time stand-in for mpi4py (absent in this image, no network).

Test infrastructure only: lets the compiled reference under oracle/_ref import.
The reference guards every compute method with ``rank == 0`` and its only real
collectives live in the FoF reader, which the oracle never calls.
"""


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        return None

    def gather(self, obj, root=0):
        return [obj]

    def bcast(self, obj, root=0):
        return obj

    def Bcast(self, buf, root=0):
        return None

    def Gatherv(self, sendbuf, recvbuf, root=0):
        import numpy as np
        dst = recvbuf[0] if isinstance(recvbuf, (list, tuple)) else recvbuf
        np.copyto(np.asarray(dst).reshape(-1)[: np.asarray(sendbuf).size], np.asarray(sendbuf).reshape(-1))


class MPI:
    COMM_WORLD = _Comm()
    INT = "INT"
    FLOAT = "FLOAT"
    DOUBLE = "DOUBLE"
