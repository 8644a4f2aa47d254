"""Mirror of the reference's ``InterpolationBasedImmersedFEA.profile_utils`` (reference profile_utils.py:9-25): every
demo imports ``profile_separate`` from here (demos/poisson.py:15), none applies it.  Same decorator; the rank comes
from mpi4py when it is importable (as in the reference) and is 0 otherwise."""
import cProfile

try:
    from mpi4py import MPI as pyMPI  # type: ignore

    _COMM_WORLD = pyMPI.COMM_WORLD
except Exception:  # pragma: no cover - depends on the environment
    pyMPI = None
    _COMM_WORLD = None


def profile_separate(filename=None, comm=_COMM_WORLD):
    """Profile the decorated function with cProfile; print the statistics, or dump them to ``filename.<rank>``."""

    def prof_decorator(f):
        def wrap_f(*args, **kwargs):
            pr = cProfile.Profile()
            pr.enable()
            result = f(*args, **kwargs)
            pr.disable()
            if filename is None:
                pr.print_stats()
            else:
                rank = comm.Get_rank() if comm is not None else 0
                pr.dump_stats(filename + ".{}".format(rank))
            return result

        return wrap_f

    return prof_decorator
