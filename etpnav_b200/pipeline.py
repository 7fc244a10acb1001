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


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned memory (first-touch:
    the staging buffers then live in that node's DRAM, next to the GPU's PCIe root).  torchrun starts its ranks unbound;
    on a two-socket HGX box GPU0-3 sit on node 0 and GPU4-7 on node 1, and a rank copying from the far socket shares
    the inter-socket link with every other rank doing the same.  Returns what was done (for bench.py's record);
    a no-op (with the reason) on single-node hosts or when sysfs does not expose the topology."""
    import os
    info = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read().strip())
        info.update(pci=bdf, numa_node=node)
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            info["reason"] = "single NUMA node (or unknown locality)"
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            info["reason"] = "no allowed CPU on the GPU's node"
            return info
        os.sched_setaffinity(0, allowed)
        info.update(bound=True, cpus=len(allowed))
    except Exception as e:   # plumbing only: never fail a run over an unreadable sysfs file
        info["reason"] = f"{type(e).__name__}: {e}"
    return info


class HostBatchStager:
    """One pinned blob, ONE host->device copy per step.

    The reference issues one ``.cuda()`` per tensor on the compute stream (ss_trainer_ETP.py:333-342, 399-417): 14 small
    copies per planner step, each with its own launch and completion latency.  Here a step's host tensors are laid out
    back to back (256-byte aligned) in one pinned uint8 blob (``pack``); ``submit`` moves the blob with a single
    ``cudaMemcpyAsync`` on a side stream into one of ``slots`` device blobs and ``get`` hands out typed VIEWS of that
    blob (no device-side copies) once the compute stream has waited on the copy's event, so the copy of step t+1 overlaps
    step t.  Tensors named in ``bf16_keys`` are stored in the blob as bf16: legitimate exactly where the consumer's first
    act is that same round-to-nearest cast (``txt_embeds``: etp_nav_inputs.txt_embeds_bf16) — it halves their bytes."""

    def __init__(self, device, slots: int = 2, bf16_keys=("txt_embeds",)):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device, priority=-1)   # copies are dispatched ahead of queued compute
        self.bf16_keys = tuple(bf16_keys)
        self.nslots = slots
        self.dev_blobs = [None] * slots
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.meta = [None] * slots
        self.head = self.tail = self.pending = 0
        self.trace = None   # set to a list to collect (start, end) timing events of every copy (diagnostics)

    def layout(self, host: dict):
        """[(key, offset, nbytes, dtype, shape)] + passthrough dict + total bytes."""
        ents, extra, off = [], {}, 0
        for k, v in host.items():
            if not torch.is_tensor(v):
                extra[k] = v
                continue
            dt = torch.bfloat16 if (k in self.bf16_keys and v.dtype == torch.float32) else v.dtype
            nbytes = v.numel() * torch.empty(0, dtype=dt).element_size()
            ents.append((k, off, nbytes, dt, tuple(v.shape)))
            off = (off + nbytes + 255) & ~255
        return ents, extra, off

    def pack(self, host: dict, blob=None):
        """Host dict -> (pinned uint8 blob, layout).  Re-packing into an existing blob reuses its pinned pages."""
        ents, extra, total = self.layout(host)
        if blob is None or blob.numel() < total:
            blob = torch.empty(total, dtype=torch.uint8).pin_memory()
        for k, off, nbytes, dt, shape in ents:
            dst = blob[off:off + nbytes].view(dt).view(shape)
            dst.copy_(host[k])      # converts fp32 -> bf16 (round to nearest even) for bf16_keys
        return blob, (ents, extra, total)

    def submit(self, blob, meta):
        if self.pending >= self.nslots:
            raise RuntimeError("HostBatchStager: all slots are in flight; call get() first")
        ents, extra, total = meta
        i = self.head
        if self.dev_blobs[i] is None or self.dev_blobs[i].numel() < total:
            self.dev_blobs[i] = torch.empty(total, dtype=torch.uint8, device=self.device)
        main = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(main)   # the slot's previous reader (enqueued before the last get()) must be done
        with torch.cuda.stream(self.stream):
            if self.trace is not None:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(self.stream)
            self.dev_blobs[i][:total].copy_(blob[:total], non_blocking=True)
            if self.trace is not None:
                t1.record(self.stream)
                self.trace.append((t0, t1))
            self.ready[i].record(self.stream)
        self.meta[i] = meta
        self.head = (i + 1) % self.nslots
        self.pending += 1

    def get(self) -> dict:
        if self.pending == 0:
            raise RuntimeError("HostBatchStager: nothing submitted")
        i = self.tail
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        ents, extra, _ = self.meta[i]
        out = dict(extra)
        d = self.dev_blobs[i]
        for k, off, nbytes, dt, shape in ents:
            out[k] = d[off:off + nbytes].view(dt).view(shape)
        self.tail = (i + 1) % self.nslots
        self.pending -= 1
        return out
