"""diffrax stand-in that does NOT integrate: `diffeqsolve` records the term and every argument of the call site and raises
`Captured`, so the generator can evaluate the reference's own joint vector field and pin the call-site arguments
(t0, t1, dt0, y0, controller tolerances).  The solver itself stays unpinned."""
LAST = {}


class Captured(Exception):
    pass


class ODETerm:
    def __init__(self, vector_field):
        self.vector_field = vector_field


class Dopri5:
    pass


class PIDController:
    def __init__(self, **kwargs):
        self.kwargs = kwargs


def diffeqsolve(term, solver, **kwargs):
    LAST.clear()
    LAST.update(term=term, solver=solver, kwargs=kwargs)
    raise Captured()
