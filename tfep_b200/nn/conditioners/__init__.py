"""Conditioner layers for autoregressive normalizing flows."""

from .made import MADE, generate_degrees
