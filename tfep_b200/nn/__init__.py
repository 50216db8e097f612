"""Neural-network layers of the MAF hot path (mirrors the layout of the reference's ``tfep.nn``)."""
