"""(T)FEP estimator and bootstrap analysis (reference tfep/analysis/__init__.py)."""

from .bootstrap import bootstrap
from .estimator import fep_estimator
