"""Stand-in for `torchdiffeq` (imported at package import by the reference's
flows/cnf.py, which is out of scope).  Calling the solvers raises."""


def _missing(*a, **k):
    raise RuntimeError("torchdiffeq is not installed; CNF layers are out of scope")


odeint = odeint_adjoint = _missing
