"""torchpdlp_b200 -- B200-native restarted-PDHG LP engine behind torchPDLP's API.

Module names mirror the reference's PDLP/ package so its call sites read the same:
``primal_dual_hybrid_gradient.pdlp_algorithm``, ``primal_dual_hybrid_gradient_step``,
``helpers``, ``enhancements``, ``spectral_casting``, ``util``.  Importing the package
does not need a GPU; building an engine does (there is no CPU fallback).
"""
__version__ = "0.1.0"

from . import synthetic  # noqa: F401


def __getattr__(name):
    import importlib
    if name in ("engine", "primal_dual_hybrid_gradient", "primal_dual_hybrid_gradient_step", "helpers",
                "enhancements", "spectral_casting", "util", "main", "_lib", "_cache", "sharded"):
        return importlib.import_module(f".{name}", __name__)
    if name == "pdlp_algorithm":
        from .primal_dual_hybrid_gradient import pdlp_algorithm
        return pdlp_algorithm
    if name == "pdlp_solver":
        from .primal_dual_hybrid_gradient import pdlp_solver
        return pdlp_solver
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
