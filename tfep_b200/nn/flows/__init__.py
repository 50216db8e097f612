"""Normalizing flow layers; every layer returns ``(y, log_det_J)`` and implements ``inverse()``."""

from .autoregressive import AutoregressiveFlow
from .maf import MAF
from .sequential import SequentialFlow
