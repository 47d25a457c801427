"""Double-buffered host → device input feed.

The reference trains with ``subgraph.to(device)`` / ``graph.to(device)`` inside the step
(src/ogbn-proteins/gat.py:109,131; src/no-sampling/run.py:539), i.e. the copy and the layer are
serialised.  At the proteins shape the per-step inputs are 2.8 GB — 50 ms of PCIe against 30 ms of
kernels — so the copy of step i+1 should run while step i computes.  :class:`HostFeed` owns a ring
of device buffers and one copy stream; ``submit`` enqueues the copies of the next step, ``take``
hands the oldest submitted set to the compute stream as :class:`bot_b200.Deferred` tensors (an
event each, so the consumer waits only where a tensor is first needed).
"""
import torch

from .functional import Deferred


class HostFeed:
    def __init__(self, device, depth=2):
        if torch.device(device).type != "cuda":
            raise RuntimeError("bot_b200.HostFeed copies to a CUDA device (bot_b200 has no CPU path)")
        self.device = torch.device(device)
        self.depth = int(depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [None] * self.depth     # per slot: list of device buffers
        self._queue = []                      # submitted, not yet taken: (slot, [(buffer, event)])
        self._next = 0

    def submit(self, *host_tensors):
        """Start copying one step's inputs (pinned host tensors).  The slot's previous contents must no
        longer be needed by work enqueued on the current stream AFTER this call; work enqueued before
        it is waited for."""
        if len(self._queue) >= self.depth:
            raise RuntimeError("HostFeed: %d steps already in flight, take() one first" % self.depth)
        slot = self._next
        self._next = (self._next + 1) % self.depth
        bufs = self._slots[slot]
        if bufs is None or len(bufs) != len(host_tensors) or any(
                b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host_tensors)):
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
            self._slots[slot] = bufs
        cur = torch.cuda.current_stream(self.device)
        reuse = torch.cuda.Event()
        reuse.record(cur)
        items = []
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(reuse)     # every consumer of this slot enqueued so far has finished
            for b, h in zip(bufs, host_tensors):
                b.copy_(h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                items.append((b, ev))
        self._queue.append((slot, items))

    def take(self, requires_grad=()):
        """Oldest submitted step as a tuple of :class:`Deferred`; ``requires_grad`` lists the positions that
        become autograd leaves when waited for."""
        if not self._queue:
            raise RuntimeError("HostFeed: nothing submitted")
        _, items = self._queue.pop(0)
        return tuple(Deferred(b.detach(), ev, requires_grad=(i in requires_grad)) for i, (b, ev) in enumerate(items))
