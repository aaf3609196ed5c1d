"""Training-step runner for the LP network: the whole step -- Network._loss (message passing, 1-N scoring, BCE),
backward, gradient all-reduce (multi-GPU), optimiser -- captured ONCE in a CUDA graph and replayed.

Why: one step enqueues ~80 libmrgnas kernels and ~150 small torch kernels; issued from Python that is 5-6 ms of
host time, the same order as the 7-8 ms of device time on one B200 and MORE than the device time once the graph
is destination-partitioned over several GPUs.  The step is static (fixed graph, fixed batch shape), so a CUDA
graph removes the host from the loop; NCCL collectives are captured with it.

The reference loop this replaces is train/mr_lp_train.py:222-246 (zero_grad, model._loss, backward, step)."""
import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, g, opt, batch_size, label_cols); loss = step(subj, rel, label).
    `subj`, `rel` [B] int64 and `label` [B, label_cols] fp32 may live on the host (pinned) or on the device;
    they are copied into the graph's static inputs.  Returns the device scalar loss of the replayed step.
    `grad_sync(params)` (optional) runs between backward and the optimiser, inside the graph.
    The optimiser must be capturable (torch.optim.Adam(..., fused=True, capturable=True))."""

    def __init__(self, model, g, opt, batch_size, label_cols, grad_sync=None, warmup=3, device=None,
                 sparse_labels=None):
        dev = device or next(model.parameters()).device
        self.model, self.g, self.opt, self.grad_sync = model, g, opt, grad_sync
        self.params = [p for p in model.parameters()]
        self.subj = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self.rel = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self.label = torch.zeros(batch_size, label_cols, dtype=torch.float32, device=dev)
        # sparse_labels = dict(num_ent=N, lbl_smooth=ls, cap=max object ids per batch, col_lo=.., col_hi=..):
        # the step starts by expanding the batch's object lists (CSR) into self.label ON THE DEVICE
        # (mrg_labels_from_csr) -- the host then sends kilobytes per step instead of the dense [B, N] matrix
        self.sparse = sparse_labels
        if sparse_labels is not None:
            self.lptr = torch.zeros(batch_size + 1, dtype=torch.int32, device=dev)
            self.lidx = torch.zeros(int(sparse_labels['cap']), dtype=torch.int32, device=dev)
        self.graph = None
        self.loss = None
        self._warmup = warmup
        # input pipeline: the NEXT batch is copied host->device on a side stream into a staging set while the
        # current step runs; the step then moves it into the graph's static inputs with device-to-device copies
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._stage = [(torch.zeros_like(self.subj), torch.zeros_like(self.rel), torch.zeros_like(self.label))
                       for _ in range(2)]
        self._staged = [None, None]      # event recorded after the staging copies
        self._consumed = [None, None]    # event recorded after the step copied the staging set out
        self._next = 0

    def _eager(self):
        if self.sparse is not None:
            from .process_data import labels_on_device
            sp = self.sparse
            labels_on_device(self.lptr, self.lidx, sp['num_ent'], sp.get('lbl_smooth', 0.0), sp.get('col_lo', 0),
                             sp.get('col_hi'), out=self.label)
        self.opt.zero_grad(set_to_none=True)
        loss = self.model._loss(self.g, self.subj, self.rel, self.label)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(self.params)
        self.opt.step()
        return loss

    def capture(self, keep_warmup_updates=False):
        """Warm up on a side stream (lazy workspaces, cudaFuncSetAttribute, NCCL channels), then capture.
        The warm-up runs `warmup` training steps on the batch currently in the static inputs; model parameters
        and buffers (BatchNorm running statistics, num_batches_tracked) and the optimiser state (moments, step
        counters) are restored afterwards, so training starts from exactly the state capture() was called in
        (`keep_warmup_updates=True` keeps them: then graph replay continues the eager steps bit-identically,
        tests/test_gpu_network_lp.py::test_graphed_train_step_replays_the_eager_step).
        No loss tensor of an earlier eager step on the default stream may still be alive: its autograd graph pins
        the parameters' AccumulateGrad nodes to the legacy stream, which cannot be used during capture."""
        import copy
        snap = None
        if not keep_warmup_updates:
            snap = ({k: v.detach().clone() for k, v in self.model.state_dict().items()},
                    copy.deepcopy(self.opt.state_dict()))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snap is not None:
            with torch.no_grad():
                for k, v in self.model.state_dict().items():     # in place: the graph captures these addresses
                    v.copy_(snap[0][k])
                had_state = len(snap[1]['state']) > 0
                if had_state:
                    cur = self.opt.state_dict()
                    for pid, st in cur['state'].items():
                        for name, val in st.items():
                            if torch.is_tensor(val):
                                val.copy_(snap[1]['state'][pid][name])
                else:            # fresh optimiser: zero the lazily created moments / step counters in place
                    for st in self.opt.state.values():
                        for val in st.values():
                            if torch.is_tensor(val):
                                val.zero_()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._eager().detach()
        return self

    def load(self, subj, rel, label):
        self.subj.copy_(subj, non_blocking=True)
        self.rel.copy_(rel, non_blocking=True)
        self.label.copy_(label, non_blocking=True)

    def load_sparse(self, subj, rel, ptr, idx):
        """Batch with sparse labels (process_data.make_batch_sparse): three small host->device copies."""
        if idx.numel() > self.lidx.numel():
            raise RuntimeError(f"batch lists {idx.numel()} objects, capacity is {self.lidx.numel()}")
        self.subj.copy_(subj, non_blocking=True)
        self.rel.copy_(rel, non_blocking=True)
        self.lptr.copy_(ptr, non_blocking=True)
        self.lidx[:idx.numel()].copy_(idx, non_blocking=True)

    def prefetch(self, subj, rel, label):
        """Start copying a (pinned host) batch into a staging set; the next __call__() without arguments runs it."""
        k = self._next
        with torch.cuda.stream(self._copy_stream):
            if self._consumed[k] is not None:
                self._copy_stream.wait_event(self._consumed[k])
            for dst, src in zip(self._stage[k], (subj, rel, label)):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged[k] = ev
        self._next = 1 - k

    def _take_staged(self):
        k = 1 - self._next              # the set filled by the most recent prefetch()
        if self._staged[k] is None:
            return False
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged[k])
        self.load(*self._stage[k])
        ev = torch.cuda.Event()
        ev.record(cur)
        self._consumed[k], self._staged[k] = ev, None
        return True

    def __call__(self, subj=None, rel=None, label=None, label_csr=None):
        if label_csr is not None:
            self.load_sparse(subj, rel, *label_csr)
        elif subj is not None:
            self.load(subj, rel, label)
        else:
            self._take_staged()
        if self.graph is None:
            return self._eager().detach()
        self.graph.replay()
        return self.loss
