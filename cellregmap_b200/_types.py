from enum import Enum


class Term(Enum):
    """Kind of model term (reference cellregmap/_types.py)."""
    FIXED = 1
    RANDOM = 2
