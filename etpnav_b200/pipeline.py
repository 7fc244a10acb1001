"""Host -> device input staging for the planner step.

In the reference every step ends with per-tensor ``.cuda()`` copies issued on the compute stream
(ss_trainer_ETP.py:333-342, 399-417): the copy of step t+1 cannot start before step t has finished.
``HostInputPrefetcher`` moves them to a side stream with two device-side slots, so the inputs of step t+1 cross
PCIe while step t computes; the compute stream only waits on an event.  Pure plumbing (pinned memory, streams,
events); no arithmetic happens here.
"""
import torch


class HostInputPrefetcher:
    """Double-buffered asynchronous staging of a dict of pinned host tensors.

        pf = HostInputPrefetcher(device)
        pf.submit(host_dict)                 # prime: copy of step 0 starts
        for t in range(steps):
            d = pf.get()                     # compute stream waits for step t's copy (event, no host sync)
            pf.submit(next_host_dict)        # step t+1's copy starts now, overlapping step t
            step(d)

    ``submit`` may be called at most once between two ``get`` calls (two slots).  Tensors that are already on the
    device are passed through untouched."""

    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [dict() for _ in range(slots)]
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.head = 0     # next slot to fill
        self.tail = 0     # next slot to hand out
        self.pending = 0

    def submit(self, host: dict):
        if self.pending >= len(self.slots):
            raise RuntimeError("HostInputPrefetcher: all slots are in flight; call get() first")
        slot = self.slots[self.head]
        main = torch.cuda.current_stream(self.device)
        # the slot was last read by the step enqueued before the previous get(): everything enqueued on the compute
        # stream so far must finish before the copy stream overwrites it
        self.stream.wait_stream(main)
        with torch.cuda.stream(self.stream):
            for k, v in host.items():
                if not torch.is_tensor(v) or v.device == self.device:
                    slot[k] = v
                    continue
                buf = slot.get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype or buf.device != self.device:
                    buf = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    slot[k] = buf
                buf.copy_(v, non_blocking=True)
            self.ready[self.head].record(self.stream)
        self.head = (self.head + 1) % len(self.slots)
        self.pending += 1

    def get(self) -> dict:
        if self.pending == 0:
            raise RuntimeError("HostInputPrefetcher: nothing submitted")
        torch.cuda.current_stream(self.device).wait_event(self.ready[self.tail])
        out = self.slots[self.tail]
        self.tail = (self.tail + 1) % len(self.slots)
        self.pending -= 1
        return dict(out)
