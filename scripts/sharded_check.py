"""Multi-GPU check + timing of the entity-sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/sharded_check.py [--rows-per-gpu 600000] [--n 128]

1. correctness: sharded ranks / BCE / logits on N ranks == the same quantities computed on rank 0 alone
   over the concatenated table (small shape, exact for integer outputs);
2. timing (BASELINE config 5 shape): TransE d=512 L1, Wikidata5M-shaped shards of `rows-per-gpu`
   entities per GPU (weak scaling in E), batch n: rank_sp_po = query-row exchange + local fused
   score+rank on the shard + int64 all-reduce; max over ranks, CUDA events."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kge_b200.sharded import ShardedKgeModel  # noqa: E402
from kge_b200 import synthetic as orc  # noqa: E402  (seeded synthetic inputs only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=600000)
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    # ---- 1. correctness on a small graph --------------------------------------------------------
    for model, D in (("transe", 64), ("complex", 64)):
        E, R, n = 4001, 7, 50
        ent, rel = orc.make_tables(model, E, R, D, sigma=0.5)
        tri = orc.make_triples(E, R, n).to(dev)
        lo, hi = ShardedKgeModel.shard_bounds(E, world, rank)
        m = ShardedKgeModel(model, ent[lo:hi].to(dev), rel.to(dev), E)
        one = ShardedKgeModel(model, ent.to(dev), rel.to(dev), E, rank=0, world=1)
        s, p, o = tri[:, 0], tri[:, 1], tri[:, 2]
        a, b = m.rank_sp_po(s, p, o), one.rank_sp_po(s, p, o)
        for x, y in zip(a, b):
            assert torch.equal(x, y), f"{model}: sharded ranks differ from single-GPU ranks"
        fa, fb = m.score_sp_po(s, p, o), one.score_sp_po(s, p, o)
        assert torch.equal(fa, fb), f"{model}: gathered logits differ"
        la, lb = float(m.loss_1vsall_bce(s, p, o)), float(one.loss_1vsall_bce(s, p, o))
        assert abs(la - lb) <= 1e-5 * abs(lb), (la, lb)
        va, ia = m.topk_sp(s, p, 10)
        vb, ib = one.topk_sp(s, p, 10)
        assert torch.equal(ia, ib) and torch.equal(va, vb)
    if rank == 0:
        print(json.dumps({"check": "sharded == single-GPU (ranks, logits, BCE, top-k)", "world": world, "ok": True}),
              flush=True)

    # ---- 1b. filtered entity ranking over the sharded table == the reference job's trace --------------
    # (golden fixtures travel with the repo; first run of this section is part of the next round's checks)
    try:
        import numpy as np
        from kge_b200.evaluate import EntityRankingEvaluator

        for model in ("complex", "transe"):
            z = np.load(os.path.join(ROOT, "tests", "golden", f"jobs_{model}.npz"))
            g = {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}
            E = g["ent"].shape[0]
            lo, hi = ShardedKgeModel.shard_bounds(E, world, rank)
            m = ShardedKgeModel(model, g["ent"][lo:hi].to(dev), g["rel"].to(dev), E)
            ev = EntityRankingEvaluator(m, E, [g["train"], g["valid"]], g["test"], batch_size=16, hits_at_k_s=(1, 3, 10),
                                        device=dev)
            met = ev.evaluate(g["valid"])
            bad = {k: (met[k[6:]], float(v)) for k, v in g.items()
                   if k.startswith("valid_") and abs(met[k[6:]] - float(v)) > 1e-6 * max(1.0, abs(float(v)))}
            if rank == 0:
                print(json.dumps({"check": f"sharded evaluation == reference job ({model})", "world": world,
                                  "ok": not bad, "mismatch": bad}), flush=True)
    except Exception as ex:  # keep the timing section running even if this new section fails
        if rank == 0:
            print(json.dumps({"check": "sharded evaluation", "world": world, "ok": False, "error": repr(ex)}), flush=True)

    # ---- 2. timing at Wikidata5M-shaped shards ---------------------------------------------------
    model, D, n = "transe", args.dim, args.n
    E = args.rows_per_gpu * world
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    shard = torch.randn((args.rows_per_gpu, D), generator=g, device=dev)
    rel = torch.randn((822, D), generator=torch.Generator(device=dev).manual_seed(7), device=dev)
    m = ShardedKgeModel(model, shard, rel, E)
    gi = torch.Generator().manual_seed(3)
    tri = torch.stack([torch.randint(0, E, (n,), generator=gi), torch.randint(0, 822, (n,), generator=gi),
                       torch.randint(0, E, (n,), generator=gi)], 1).to(dev)
    s, p, o = tri[:, 0], tri[:, 1], tri[:, 2]
    for _ in range(3):
        m.rank_sp_po(s, p, o)
    dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
    for a, b in ev:
        a.record()
        m.rank_sp_po(s, p, o)
        b.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.iters], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"TransE d={D} L1 entity-sharded rank_sp_po, {args.rows_per_gpu} rows/GPU, "
                                      f"E={E}, n={n}", "n_gpus": world, "ms_per_call": float(ms),
                          "triples_per_s": 2.0 * n * E / (float(ms) * 1e-3)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
