"""ddo_b200 -- B200-native MDD compilation behind ddo's solver API (MISP and MAX2SAT device models).

Python host-side mirror of the reference's interface for the hot path; every class / method keeps the reference's name and argument
meaning (ddo/src/abstraction/{dp,mdd,solver,heuristics}.rs, ddo/examples/misp/main.rs).  All compute goes through the C ABI of
include/ddo_b200.h into hand-written sm_100a kernels; there is no CPU fallback.
"""
from .api import (CompilationType, Completion, CutoffOccurred, Decision, DivBy, FixedWidth, Times, GpuMdd, LAST_EXACT_LAYER, FRONTIER, Max2Sat, Misp, NbUnassignedWidth,
                  ParNoCachingSolverLel, ParNoCachingSolverFc, DefaultSolver, SubProblem, device_count, kernel_launches)
from .instances import (Max2SatInstance, MispInstance, TsptwInstance, gnp, parse_dimacs, parse_tsptw, parse_wcnf, random_max2sat, read_dimacs, read_tsptw,
                        read_wcnf)

__all__ = ["Times", "DivBy", "CompilationType", "Completion", "CutoffOccurred", "Decision", "FixedWidth", "GpuMdd", "LAST_EXACT_LAYER", "FRONTIER", "Max2Sat", "Misp",
           "NbUnassignedWidth", "ParNoCachingSolverLel", "ParNoCachingSolverFc", "DefaultSolver", "SubProblem", "device_count", "kernel_launches", "MispInstance", "gnp",
           "parse_dimacs", "read_dimacs", "Max2SatInstance", "parse_wcnf", "random_max2sat", "read_wcnf", "TsptwInstance", "parse_tsptw", "read_tsptw"]
