"""Single-rank stand-in for the few mpi4py calls the reference makes.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): lets the unmodified reference under
/root/reference/src be imported in a container without MPI, so that golden vectors can be
generated from it (tests/golden/make_golden.py).  Never imported by the product package.
Calls covered: core/mgrit.py:13,124-141,428-432,637,649; core/split.py:7; core/at_mgrit.py:11.
"""


class _Request:
    @staticmethod
    def Waitall(requests):
        return None


class _Comm:
    rank = 0
    size = 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def barrier(self):
        return None

    def gather(self, value, root=0):
        return [value]

    def bcast(self, value, root=0):
        return value

    def allgather(self, value):
        return [value]

    def Split(self, color=0, key=0):
        return self


class MPI:
    Comm = _Comm
    COMM_WORLD = _Comm()
    COMM_NULL = None
    Request = _Request
    UNDEFINED = -32766
