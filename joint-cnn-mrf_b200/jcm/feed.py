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
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pending = None

    def submit(self, *host_tensors):
        if self._pending is not None:
            raise RuntimeError('DeviceFeed.submit: the previous batch was not taken')
        for t in host_tensors:
            if not t.is_pinned():
                raise ValueError('DeviceFeed needs pinned host tensors (torch.Tensor.pin_memory())')
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev, ev)

    def take(self):
        if self._pending is None:
            raise RuntimeError('DeviceFeed.take: nothing was submitted')
        dev, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)      # the tensors were allocated on the feed's stream but are consumed on `cur`
        return dev
