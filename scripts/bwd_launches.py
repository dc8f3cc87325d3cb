"""One call of the native 1vsAll backward per model at the headline shape, for an ncu launch list
(ncu --metrics gpu__time_duration.sum ... python scripts/bwd_launches.py complex transe)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kge_b200 import engine, synthetic  # noqa: E402

E, R, n, D = 14541, 237, 1024, 512
for model in (sys.argv[1:] or ["complex"]):
    ent, rel = (t.cuda() for t in synthetic.make_tables(model, E, R, D, sigma=0.5))
    tri = synthetic.make_triples(E, R, n).cuda()
    engine.train_1vsall_backward(model, ent, rel, tri, "bce", 0.0, 1.0)
    torch.cuda.synchronize()
