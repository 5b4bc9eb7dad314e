"""Host -> device input feed for the training / inference loop (the reference feeds numpy batches through `feed_dict`,
main.py:640-648, a synchronous host->device copy inside `sess.run`).  Here the copy of batch i+1 from pinned host memory
runs on its own CUDA stream while batch i is being computed, so the PCIe transfer (276 MB per 64-image batch) is hidden
behind the step instead of being added to it."""
import torch


class DeviceFeed:
    """Double-buffered asynchronous host->device copies.

        feed = DeviceFeed(device)
        feed.submit(x_host, y_host)          # pinned host tensors; starts the copy on the feed's stream
        x, y = feed.take()                   # makes the current stream wait for that copy, returns the device tensors

    The device tensors live in two persistent buffer sets used alternately (no allocation per batch): the tensors returned by
    take() are overwritten by the second submit() after it, so a step must have been ENQUEUED (not finished) on the current stream
    before the submit after next - the natural order of a training loop (take, submit next, step)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pending = None
        self._bufs = [None, None]
        self._n = 0

    def submit(self, *host_tensors):
        if self._pending is not None:
            raise RuntimeError('DeviceFeed.submit: the previous batch was not taken')
        for t in host_tensors:
            if not t.is_pinned():
                raise ValueError('DeviceFeed needs pinned host tensors (torch.Tensor.pin_memory())')
        slot = self._n & 1
        self._n += 1
        bufs = self._bufs[slot]
        if bufs is None or len(bufs) != len(host_tensors) or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, host_tensors)):
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
            self._bufs[slot] = bufs
        # this buffer set was last read by the step enqueued two submits ago: the copy waits for everything enqueued so far
        guard = torch.cuda.Event()
        guard.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(guard)
        with torch.cuda.stream(self.stream):
            for b, t in zip(bufs, host_tensors):
                b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (bufs, ev)

    def take(self):
        if self._pending is None:
            raise RuntimeError('DeviceFeed.take: nothing was submitted')
        dev, ev = self._pending
        self._pending = None
        torch.cuda.current_stream(self.device).wait_event(ev)
        return list(dev)
