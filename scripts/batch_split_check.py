"""Multi-GPU check of `user.b200_batch_split` (replicas + batch split, SURVEY 8e "small tables"), under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        scripts/batch_split_check.py

Every rank trains two epochs of the fused 1vsAll / KvsAll / negative-sampling job plugins on its GPU with the option on
(its rows of every batch, ncclAllReduce of the dense table gradients), rank 0 also trains the same job alone with the
option off: loss trajectories and final tables must agree, and all replicas must hold identical tables."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

E, R, D = 301, 6, 64
CASES = {
    "1vsAll": dict(model="complex", train_type="1vsAll", loss="kl", batch_size=96, job_class="B200TrainingJob1vsAll"),
    "1vsAll-transe": dict(model="transe", train_type="1vsAll", loss="bce", batch_size=96, job_class="B200TrainingJob1vsAll"),
    "KvsAll": dict(model="distmult", train_type="KvsAll", loss="bce", batch_size=48, job_class="B200TrainingJobKvsAll"),
    "negative_sampling": dict(model="complex", train_type="negative_sampling", loss="bce", batch_size=96,
                              job_class="B200TrainingJobNegativeSampling",
                              extra={"negative_sampling.num_samples.s": 20, "negative_sampling.num_samples.o": 30}),
}


def train(case, split, device):
    import jobs_util as ju

    kw = dict(CASES[case])
    model = kw.pop("model")
    extra = dict(kw.pop("extra", {}))
    extra["user.b200_batch_split"] = split
    splits = ju.synthetic_splits(E, R, 700, 20, 20)
    torch.manual_seed(0)
    init = ju.make_job(model, E, R, D, splits, **{k: v for k, v in kw.items() if k != "job_class"})
    job = ju.make_job("b200_" + model, E, R, D, splits, device=device, forward_only=False, extra=extra, **kw)
    job.model.b200_backward = "native"
    ju.copy_tables(init, job)
    losses = []
    for ep in range(2):
        job.epoch += 1
        if job.loader is None:
            job._prepare()
        ju.seed_all(10 + ep)
        losses.append(job.run_epoch()["avg_loss"])
    return (losses, job.model.get_s_embedder()._embeddings.weight.detach().clone(),
            job.model.get_p_embedder()._embeddings.weight.detach().clone())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {}
    for case in CASES:
        losses, ent, rel = train(case, True, f"cuda:{local}")
        lo, hi = ent.clone(), ent.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), f"{case}: replicas diverged"
        if rank == 0:
            ref_losses, ref_ent, ref_rel = train(case, False, f"cuda:{local}")
            for a, b in zip(losses, ref_losses):
                assert abs(a - b) <= 1e-4 * abs(b), (case, losses, ref_losses)
            # Frobenius norm: Adagrad turns a gradient element that is round-off around zero into a +-lr step, so single
            # elements may differ by 2 lr between two summation orders (the L1 models' sign gradients make that common)
            err = max(float((ent - ref_ent).norm() / ref_ent.norm()), float((rel - ref_rel).norm() / ref_rel.norm()))
            # noise floor: the same single-process job twice (atomic accumulation order in the unfold kernels)
            _, ent2, rel2 = train(case, False, f"cuda:{local}")
            noise = max(float((ent2 - ref_ent).norm() / ref_ent.norm()), float((rel2 - ref_rel).norm() / ref_rel.norm()))
            report[case] = {"avg_loss": losses, "single_process": ref_losses, "table_rel_err": err,
                            "single_process_run_to_run": noise,
                            "elements_differing_by_more_than_1e-3": int(((ent - ref_ent).abs() > 1e-3).sum()),
                            "elements": ent.numel()}
            assert err <= max(1e-3, 4 * noise), (case, report[case])
        dist.barrier()
    if rank == 0:
        print(json.dumps({"check": "batch split over N ranks == single-process job", "world": world, "cases": report}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
