"""One evaluation of a right-hand side for profiling: python tools/one_eval.py {div|f256|drift} [n_mol]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import perturb_  # noqa: E402
from thermodynamic_interpolation_b200.ambient.models.cpainn import cPaiNN  # noqa: E402
from thermodynamic_interpolation_b200.batch import synthetic_ambient_batch  # noqa: E402

what = sys.argv[1]
F, n = (256, 25) if what == "f256" else (128, 9)
n_mol = int(sys.argv[2]) if len(sys.argv) > 2 else (512 if what == "f256" else 4096)
torch.manual_seed(0)
model = perturb_(cPaiNN(n_features=F, score_layers=5, temp_length=100), 1).eval().to("cuda:0")
if len(sys.argv) > 3:
    model.set_math(int(sys.argv[3]))
mb = synthetic_ambient_batch(n_mol, n, seed=100).to("cuda:0")
eng = model.engine()
pb = eng.prepare(mb)
x = mb.x0.contiguous()
if what == "div":
    eng.drift_div(pb, x, 0.3)
else:
    eng.drift(pb, x, 0.3)
torch.cuda.synchronize()
eng.status()
