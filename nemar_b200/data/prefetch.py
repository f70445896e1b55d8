"""Device input pipeline: pinned, double-buffered host->device staging of the {'A','B'} batches.

The reference moves every batch with a blocking `.to(device)` from pageable memory inside `set_input`
(models/nemar_model.py:151-159; the loader is built without pin_memory, data/__init__.py:75-79), so the copy of
batch i+1 cannot start before step i has been launched AND finished copying.  Here a batch is staged into one of
`depth` pinned host buffers and copied on a side stream into one of `depth` device buffers while the previous step
computes; the consumer only orders its stream behind the copy's event (`NEMARModel.set_input` does that when the
batch dict carries `_ready`).  Works for any iterable of dicts with tensor values; non-tensor values pass through.

    for data in DevicePrefetcher(dataset, device):      # train.py
        model.set_input(data); model.optimize_parameters()
"""
import torch


class DevicePrefetcher:
    def __init__(self, loader, device, depth=2, keys=("A", "B")):
        self.loader, self.device, self.depth, self.keys = loader, torch.device(device), max(2, int(depth)), tuple(keys)
        self.stream = None
        self.slots = [None] * self.depth

    def __len__(self):
        return len(self.loader)

    def _slot(self, i, batch):
        """buffers of slot i, (re)allocated when the batch shape changes (a short last batch)"""
        s = self.slots[i]
        shapes = {k: (tuple(batch[k].shape), batch[k].dtype) for k in self.keys}
        if s is None or s["shapes"] != shapes:
            # New device buffers are FIRST WRITTEN on the side stream.  The caching allocator may hand out memory that
            # pending work on the compute stream still reads or writes (a freed temporary of a step that has not run
            # yet), so the side stream is ordered behind everything enqueued so far, and the buffers are taken from the
            # side stream's own pool.  (Rare: the first `depth` batches and shape changes.)
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            if s is not None and s["busy"] is not None:
                s["busy"].synchronize()
            s = {"shapes": shapes, "pin": {}, "dev": {}, "free": None, "busy": None}
            with torch.cuda.stream(self.stream):
                for k in self.keys:
                    # pinned staging only for pageable sources (cudaHostAlloc costs ~10 ms per buffer)
                    s["pin"][k] = None if batch[k].is_pinned() else torch.empty(batch[k].shape, dtype=batch[k].dtype, pin_memory=True)
                    s["dev"][k] = torch.empty(batch[k].shape, dtype=batch[k].dtype, device=self.device)
                    s["dev"][k].record_stream(torch.cuda.current_stream(self.device))
            self.slots[i] = s
        return s

    def _stage(self, i, batch):
        """host batch -> slot i: pinned staging (skipped when the loader already pins), async H2D on the side stream"""
        s = self._slot(i, batch)
        if s["free"] is not None:
            self.stream.wait_event(s["free"])            # the step that read this slot's device buffers has finished
        out = dict(batch)
        with torch.cuda.stream(self.stream):
            for k in self.keys:
                src = batch[k]
                if not src.is_pinned():
                    if s["pin"][k] is None:
                        s["pin"][k] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                    if s["busy"] is not None:
                        s["busy"].synchronize()          # the previous H2D out of this pinned buffer has completed
                    s["pin"][k].copy_(src)
                    src = s["pin"][k]
                s["dev"][k].copy_(src, non_blocking=True)
                out[k] = s["dev"][k]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        s["busy"] = ev
        out["_ready"] = ev
        out["_slot"] = i
        return out

    def __iter__(self):
        if self.device.type != "cuda":
            yield from self.loader
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream(self.device)
        it = iter(self.loader)
        i = 0
        try:
            cur = self._stage(0, next(it))
        except StopIteration:
            return
        while True:
            try:
                nxt_host = next(it)
            except StopIteration:
                nxt_host = None
            nxt = None
            if nxt_host is not None:
                nxt = self._stage((i + 1) % self.depth, nxt_host)     # overlaps the step the consumer runs on `cur`
            yield cur
            # everything the consumer enqueued for `cur` is on its stream now: its slot is free once that work is done
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self.slots[cur["_slot"]]["free"] = done
            if nxt is None:
                return
            cur = nxt
            i += 1
