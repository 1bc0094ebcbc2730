"""cellregmap_b200 -- B200 (sm_100a) implementation of CellRegMap's per-variant scans behind the
reference's API (reference cellregmap/__init__.py:1-20)."""
from ._types import Term
from ._cellregmap import (
    CellRegMap,
    compute_maf,
    estimate_betas,
    get_L_values,
    lrt_pvalues,
    run_association,
    run_association_fast,
    run_interaction,
)

__version__ = "0.1.0"

__all__ = [
    "__version__",
    "CellRegMap",
    "run_association",
    "run_association_fast",
    "run_interaction",
    "estimate_betas",
    "get_L_values",
    "compute_maf",
    "lrt_pvalues",
    "Term",
]
