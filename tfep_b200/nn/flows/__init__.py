"""Normalizing flow layers; every layer returns ``(y, log_det_J)`` and implements ``inverse()``."""

from .autoregressive import AutoregressiveFlow
from .centroid import CenteredCentroidFlow
from .maf import MAF
from .oriented import OrientedFlow
from .partial import PartialFlow
from .sequential import SequentialFlow
