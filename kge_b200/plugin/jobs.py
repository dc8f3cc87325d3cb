"""Job plugins: thin subclasses of the reference's own training jobs whose `_process_subbatch` calls the FUSED
entry points of libb200kge (score + loss in one launch sequence, scores never reach HBM).

Selected through the reference's job factory (kge/job/train.py:127-137: `init_from(config.get("<type>.class_name"),
config.modules(), ...)`), i.e. in a LibKGE config:

    modules: [kge.job, kge.model, kge.model.embedder, kge_b200.plugin]
    model: b200_complex
    1vsAll.class_name: B200TrainingJob1vsAll
    KvsAll.class_name: B200TrainingJobKvsAll
    negative_sampling.class_name: B200TrainingJobNegativeSampling

Everything else of the job (data loading, collate, sub-batching, trace entries, penalties, optimizer, hooks,
checkpoints) is the reference's code, unchanged.  Whenever a fused form is not available for the configured
combination (a non-b200 model, a loss other than bce / kl, dropout active, ...) the method falls through to
the reference implementation, which then still reaches the kernels through `model.score_*`.
"""
from __future__ import annotations

import time

import torch

from kge.job.train_1vsAll import TrainingJob1vsAll
from kge.job.train_KvsAll import TrainingJobKvsAll
from kge.job.train_negative_sampling import TrainingJobNegativeSampling
from kge.job import Job
from kge.util.loss import BCEWithLogitsKgeLoss, KLDivWithSoftmaxKgeLoss

S, P, O = 0, 1, 2
SLOT_STR = ["s", "p", "o"]


def _fused_loss_kind(loss):
    """("bce"|"kl", offset) if the job's KgeLoss has a fused form, else None (loss.py:139-213)."""
    if type(loss) is BCEWithLogitsKgeLoss and loss._bce_type is None:
        return "bce", float(loss._offset)
    if type(loss) is KLDivWithSoftmaxKgeLoss:
        return "kl", 0.0
    return None


def _fused_model(model):
    """The b200 model if its fused entry points can read the tables in place, else None."""
    if getattr(model, "_b200_name", None) is None or not hasattr(model, "b200_fusable"):
        return None
    return model if model.b200_fusable() else None


def _user_option(config, key, default=None):
    try:
        return config.get("user." + key)
    except KeyError:
        return default


class _BatchSplit:
    """Replicas + batch split over the processes of a torch.distributed group ("small tables": every GPU
    holds the whole tables, scores its share of the batch's rows against them, and the dense table gradients are
    all-reduced before the optimizer step; no collective on the forward data path).

    Enabled by `user.b200_batch_split: true` when a process group with more than one rank is initialised (one process
    per GPU, `job.device: cuda:<LOCAL_RANK>`, one output folder per rank, the same random seed on every rank so that
    all ranks draw the same batches).  Rank r takes rows [r*B/W, (r+1)*B/W) of every batch; the per-row losses are
    divided by the size of the WHOLE batch (as for sub-batches, train_1vsAll.py:65), so the all-reduced gradients and
    avg_loss are those of the single-process job.  Penalties and the optimizer step run identically on every rank
    (kge/job/train.py:411-470), which keeps the replicas in step without a broadcast."""

    def _b200_ranks(self):
        r = getattr(self, "_b200_rank_world", None)
        if r is None:
            r = (0, 1)
            if _user_option(self.config, "b200_batch_split", False):
                import torch.distributed as dist

                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    r = (dist.get_rank(), dist.get_world_size())
            self._b200_rank_world = r
        return r

    def _b200_my_rows(self, subbatch_slice, batch_size):
        """This rank's part of a sub-batch slice, or None if it has none."""
        rank, world = self._b200_ranks()
        if world == 1:
            return subbatch_slice
        lo, hi = rank * batch_size // world, (rank + 1) * batch_size // world
        start = max(subbatch_slice.start or 0, lo)
        stop = min(batch_size if subbatch_slice.stop is None else subbatch_slice.stop, hi)
        return slice(start, stop) if stop > start else None

    def _process_batch(self, batch_index, batch):
        result = super()._process_batch(batch_index, batch)
        rank, world = self._b200_ranks()
        if world == 1:
            return result
        import torch.distributed as dist

        if not self.is_forward_only:
            for p in self.model.parameters():
                if not p.requires_grad:
                    continue
                if p.grad is None:                      # a rank without rows (batch smaller than the group)
                    p.grad = torch.zeros_like(p)
                if p.grad.is_sparse:
                    raise NotImplementedError("user.b200_batch_split all-reduces dense gradients (sparse: False)")
                dist.all_reduce(p.grad)
        total = torch.tensor([result.avg_loss], dtype=torch.float64, device=self.device)
        dist.all_reduce(total)
        result.avg_loss = float(total.item())
        return result


class B200TrainingJob1vsAll(_BatchSplit, TrainingJob1vsAll):
    """`TrainingJob1vsAll` (train_1vsAll.py:10-82) with the sub-batch step as ONE fused call:
    (loss(score_sp, o) + loss(score_po, s)) / batch_size, both directions stacked into one problem."""

    def __init__(self, config, dataset, parent_job=None, model=None, forward_only=False):
        super().__init__(config, dataset, parent_job, model=model, forward_only=forward_only)
        if self.__class__ == B200TrainingJob1vsAll:
            for f in Job.job_created_hooks:
                f(self)

    def _process_subbatch(self, batch_index, batch, subbatch_slice, result):
        subbatch_slice = self._b200_my_rows(subbatch_slice, result.size)
        if subbatch_slice is None:
            return
        model, kind = _fused_model(self.model), _fused_loss_kind(self.loss)
        if model is None or kind is None:
            return super()._process_subbatch(batch_index, batch, subbatch_slice, result)
        batch_size = result.size

        host = batch["triples"][subbatch_slice]
        if (self.is_forward_only and not host.is_cuda and host.dtype == torch.int64 and host.is_contiguous()
                and len(host) > 0):
            # forward only: batch copy, kernels and the scalar read-back in ONE library call (no torch ops in between)
            result.forward_time -= time.time()
            value = model.loss_1vsall_host(host, kind[0], kind[1])
            if len(host) != batch_size:
                value *= len(host) / batch_size
            result.avg_loss += value
            result.forward_time += time.time()
            return

        result.prepare_time -= time.time()
        triples = host.to(self.device, non_blocking=True)
        result.prepare_time += time.time()

        result.forward_time -= time.time()
        # sum over both directions and all rows of the sub-batch, divided by the sub-batch size by the kernel's
        # finaliser; the reference divides by the size of the whole batch (train_1vsAll.py:65,76)
        loss_value = model.loss_1vsall(triples, kind[0], kind[1], need_grad=not self.is_forward_only)
        if len(triples) != batch_size:
            loss_value = loss_value * (len(triples) / batch_size)
        result.avg_loss += loss_value.item()
        result.forward_time += time.time()

        result.backward_time -= time.time()
        if not self.is_forward_only:
            loss_value.backward()
        result.backward_time += time.time()


class B200TrainingJobKvsAll(_BatchSplit, TrainingJobKvsAll):
    """`TrainingJobKvsAll` (train_KvsAll.py:205-294): per query type one fused score+loss call that consumes the
    batch's label coordinates as CSR (no dense [n, E] label matrix: job/util.py:32-60 + `.to_dense()`,
    train_KvsAll.py:242-266 are not executed)."""

    def __init__(self, config, dataset, parent_job=None, model=None, forward_only=False):
        super().__init__(config, dataset, parent_job, model=model, forward_only=forward_only)
        if self.__class__ == B200TrainingJobKvsAll:
            for f in Job.job_created_hooks:
                f(self)

    def _process_subbatch(self, batch_index, batch, subbatch_slice, result):
        subbatch_slice = self._b200_my_rows(subbatch_slice, result.size)
        if subbatch_slice is None:
            return
        model, kind = _fused_model(self.model), _fused_loss_kind(self.loss)
        qtypes = [q for q in self.query_types]
        if (model is None or kind is None or "s_o" in qtypes or not model.b200_csr_labels_ok(self.label_smoothing)
                or (not self.is_forward_only and not model.b200_kvsall_native_backward_ok())):
            return super()._process_subbatch(batch_index, batch, subbatch_slice, result)
        batch_size = result.size

        result.prepare_time -= time.time()
        queries = batch["queries"][subbatch_slice].to(self.device)
        qt = batch["query_type_indexes"][subbatch_slice].to(self.device)
        coords = batch["label_coords"]                       # [nnz, 2] int, rows ascending (collate order)
        start, stop = subbatch_slice.start or 0, subbatch_slice.stop
        rows = coords[:, 0].long()
        if start != 0 or stop < batch_size:
            keep = (rows >= start) & (rows < stop)
            rows, cols = rows[keep] - start, coords[keep, 1].long()
        else:
            cols = coords[:, 1].long()
        n_sub = len(queries)
        result.prepare_time += time.time()

        for query_type_index, query_type in enumerate(self.query_types):
            result.prepare_time -= time.time()
            examples = (qt == query_type_index).nonzero(as_tuple=False).view(-1)
            if len(examples) == 0:
                result.prepare_time += time.time()
                continue
            # CSR of the selected rows: counts per selected row, columns in row order
            sel = torch.zeros(n_sub, dtype=torch.bool, device=self.device)
            sel[examples] = True
            keep = sel[rows]
            counts = torch.bincount(rows[keep], minlength=n_sub)[examples]
            offsets = torch.zeros(len(examples) + 1, dtype=torch.int64, device=self.device)
            torch.cumsum(counts, 0, out=offsets[1:])
            ccols = cols[keep]
            result.prepare_time += time.time()

            result.forward_time -= time.time()
            # sp_ queries are (s, p) pairs, _po queries are (p, o) pairs (indexing.py:197-235)
            if query_type == "sp_":
                combine, ent_idx, rel_idx = "sp_", queries[examples, 0], queries[examples, 1]
            else:
                combine, ent_idx, rel_idx = "_po", queries[examples, 1], queries[examples, 0]
            if self.is_forward_only:
                loss_value = model.loss_kvsall(combine, ent_idx, rel_idx, offsets, ccols, kind[0], kind[1],
                                               self.label_smoothing) / batch_size
            else:
                loss_value = model.loss_kvsall_train(combine, ent_idx, rel_idx, offsets, ccols, kind[0], kind[1],
                                                     self.label_smoothing, batch_size)
            result.avg_loss += loss_value.item()
            result.forward_time += time.time()
            result.backward_time -= time.time()
            if not self.is_forward_only:
                loss_value.backward()
            result.backward_time += time.time()


class B200TrainingJobNegativeSampling(_BatchSplit, TrainingJobNegativeSampling):
    """`TrainingJobNegativeSampling` (train_negative_sampling.py:103-164): per slot ONE kernel gathers the sampled
    rows and scores them, with the positive triple in column 0 — neither `[n*K, D]` gathers (`triple`
    implementation, sampler.py:294-305) nor scoring against all unique targets (`batch`, :306-339).

    With `user.b200_device_sampling: true` (LibKGE's free-form `user.*` option space) and the default sampler
    settings (uniform, not shared, no filtering) the negatives are also DRAWN on the device (Philox, keyed by the
    torch seed, counter = batch / slot): the DataLoader workers only slice the triples, and no [n, K] id tensors
    travel host -> device (KgeUniformSampler._sample, sampler.py:588-596, is a CPU torch.randint)."""

    def __init__(self, config, dataset, parent_job=None, model=None, forward_only=False):
        super().__init__(config, dataset, parent_job, model=model, forward_only=forward_only)
        try:
            want = bool(config.get("user.b200_device_sampling"))
        except KeyError:
            want = False
        sm = self._sampler
        self._device_sampling = bool(
            want and type(sm).__name__ == "KgeUniformSampler" and not sm.shared and not any(sm.filter_positives))
        self._sample_calls = 0
        if self.__class__ == B200TrainingJobNegativeSampling:
            for f in Job.job_created_hooks:
                f(self)

    def _get_collate_fun(self):
        if not self._device_sampling:
            return super()._get_collate_fun()

        def collate(batch):          # the triples only: negatives are drawn on the device
            return {"triples": self.dataset.split(self.train_split)[batch, :].long(), "negative_samples": []}
        return collate

    def _device_negatives(self, n, slot, batch_index):
        from .. import engine

        sm = self._sampler
        # one independent stream per (epoch, batch, slot); the key follows torch.manual_seed
        offset = ((self.epoch * (1 << 24) + batch_index) << 2) | slot
        self._sample_calls += 1
        return engine.sample_uniform(n, int(sm.num_samples[slot]), int(sm.vocabulary_size[slot]),
                                     torch.initial_seed(), offset, self.device)

    def _process_subbatch(self, batch_index, batch, subbatch_slice, result):
        subbatch_slice = self._b200_my_rows(subbatch_slice, result.size)
        if subbatch_slice is None:
            return
        model = _fused_model(self.model)
        kind = _fused_loss_kind(self.loss)
        slots = [sl for sl in (S, P, O) if self._sampler.num_samples[sl] > 0]
        trainable = (model is not None and kind is not None and kind[0] == "bce"
                     and all(model.b200_ns_native_backward_ok(sl) for sl in slots))
        if model is None or (not self.is_forward_only and not trainable):
            if self._device_sampling:
                raise NotImplementedError("user.b200_device_sampling needs a b200_* model whose slots the fused "
                                          "gradient kernel covers (S / O slots, bce)")
            return super()._process_subbatch(batch_index, batch, subbatch_slice, result)
        batch_size = result.size
        result.prepare_time -= time.time()
        triples = batch["triples"][subbatch_slice]
        negs = batch["negative_samples"]
        subbatch_size = len(triples)
        labels = batch["labels"]
        result.prepare_time += time.time()

        for slot in [S, P, O]:
            num_samples = self._sampler.num_samples[slot]
            if num_samples <= 0:
                continue
            result.prepare_time -= time.time()
            if self._device_sampling:
                if negs == [] or len(negs) < 3:
                    negs = batch["negative_samples"] = [None, None, None]
                if negs[slot] is None:          # drawn once per batch and slot, sliced per sub-batch
                    negs[slot] = self._device_negatives(batch_size, slot, batch_index)
                negatives = negs[slot][subbatch_slice]
            else:
                negatives = negs[slot].samples(subbatch_slice if (subbatch_size != batch_size) else None)
            result.prepare_time += time.time()

            result.forward_time -= time.time()
            if not self.is_forward_only:
                # training: forward + the fused NS gradient kernel behind one autograd node
                loss_value = model.loss_negatives(triples, negatives.to(self.device), slot, kind[1], batch_size)
                result.avg_loss += loss_value.item()
                result.forward_time += time.time()
                result.backward_time -= time.time()
                loss_value.backward()
                result.backward_time += time.time()
                continue
            scores = model.score_negatives(triples, negatives.to(self.device), slot)      # [n, 1+K], positive first
            if kind is not None and kind[0] == "bce":
                # labels are 1 in column 0 and 0 elsewhere (train_negative_sampling.py:128-137): index labels
                lab = torch.zeros(subbatch_size, dtype=torch.int64, device=self.device)
                loss_value = model.loss_dense(scores, lab, "bce", kind[1]) / batch_size
            else:
                if labels[slot] is None or labels[slot].shape != (subbatch_size, 1 + num_samples):
                    labels[slot] = torch.zeros((subbatch_size, 1 + num_samples), device=self.device)
                    labels[slot][:, 0] = 1
                loss_value = self.loss(scores, labels[slot], num_negatives=num_samples) / batch_size
            result.avg_loss += loss_value.item()
            result.forward_time += time.time()
