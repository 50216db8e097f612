"""Host-side helpers (reference: tfep/utils/misc.py)."""

import numpy as np
import torch


def ensure_tensor_sequence(x, dtype=None):
    """Sequences (not strings / scalars) become tensors, without a copy when possible.

    Reference: tfep/utils/misc.py:158-179.
    """
    if not np.isscalar(x):
        try:
            x = torch.as_tensor(x, dtype=dtype)
        except (TypeError, RuntimeError):
            pass
    return x
