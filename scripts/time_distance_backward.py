"""Timing of the distance-family 1vsAll backward (grad_distance.cu) on FB15k-237-shaped tables: one line of json per
(model, l_norm).  python scripts/time_distance_backward.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kge_b200 import engine, synthetic


def main():
    E, R, n = 14541, 237, 1024
    for model, D, ln in (("transe", 512, 1.0), ("transe", 512, 2.0), ("rotate", 512, 1.0), ("complex", 512, 1.0)):
        ent, rel = (t.cuda() for t in synthetic.make_tables(model, E, R, D, sigma=0.5))
        tri = synthetic.make_triples(E, R, n).cuda()
        for _ in range(3):
            engine.train_1vsall_backward(model, ent, rel, tri, "kl", 0.0, ln)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            engine.train_1vsall_backward(model, ent, rel, tri, "kl", 0.0, ln)
        b.record()
        torch.cuda.synchronize()
        print(json.dumps({"model": model, "l_norm": ln, "D": D, "E": E, "n": n, "backward_ms": a.elapsed_time(b) / 5}))


if __name__ == "__main__":
    main()
