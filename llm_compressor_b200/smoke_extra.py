"""Additional smoke checks, filled in as hot-path stages land (Hessian, solver, masks)."""


def run(dev):
    return
