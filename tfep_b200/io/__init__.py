"""On-disk formats shared with the reference (tfep/io): per-sample quantities logged during training / evaluation."""

from .log import TFEPLogger
