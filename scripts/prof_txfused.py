"""One SOS layer and one Moebius layer of cfg3 (D = 300, batch 262144), forward + backward with the fused transformer
epilogue: the launches ncu captures for profiles/ (python scripts/prof_txfused.py [sos|moebius|spline])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfep_b200.nn.conditioners.made import generate_degrees  # noqa: E402
from tfep_b200.nn.flows import MAF  # noqa: E402
from tfep_b200.nn.transformers import MoebiusTransformer, SOSPolynomialTransformer  # noqa: E402

dev = 'cuda:0'
which = sys.argv[1] if len(sys.argv) > 1 else 'both'
torch.manual_seed(4321)
layers = []
if which in ('sos', 'both'):
    layers.append(MAF(generate_degrees(300), SOSPolynomialTransformer(2), initialize_identity=False, precision='bf16').to(dev))
if which in ('moebius', 'both'):
    layers.append(MAF(generate_degrees(300, repeats=3), MoebiusTransformer(dimension=3), initialize_identity=False,
                      precision='bf16').to(dev))
if which == 'spline':
    # one cfg2 layer (D = 66, MADE 66-328-328-2112 padded, batch 65536) with the fused 8-bin spline epilogue
    import math
    from tfep_b200.nn.transformers import NeuralSplineTransformer
    lim = torch.full((66,), math.pi)
    layers.append(MAF(generate_degrees(66), NeuralSplineTransformer(x0=-lim, xf=lim, n_bins=8, circular=True),
                      initialize_identity=False, precision='bf16').to(dev))
    x = (torch.rand(65536, 66, device=dev) * 2 - 1) * math.pi * 0.999
else:
    x = torch.randn(262144, 300, device=dev)
for it in range(2):
    for maf in layers:
        maf.zero_grad(set_to_none=True)
        y, ld = maf(x)
        (y.sum() + (ld.sum() if ld.requires_grad else 0.0)).backward()
torch.cuda.synchronize()
print('ok')
