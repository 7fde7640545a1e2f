"""Pinned-host ring buffer feeding decoded uint8 tiles to the GPU on a side CUDA stream.

Replaces the reference's DataLoader(pin_memory=True) -> ``batch.to(device)`` of *float32 CHW* images
(/root/reference/data_loading/data_module.py:16-29, pytorch_loader.py:163-171): tiles cross PCIe as the uint8 HWC bytes
cv2 decoded (3 B / pixel instead of 12), the copy of batch i+1 overlaps the compute of batch i, and the
Normalize + HWC->NHWC-bf16 conversion is one kernel on the device (xv2_normalize_tiles).

    ring = TileRing(batch, h, w, post=False, device=...)
    slot = ring.host(i)            # numpy views {"tiles", ["tiles_post"], "mask"} for the decode threads to fill
    ring.submit(i)                 # async H2D on the side stream
    batch = ring.acquire(i)        # compute stream waits for the copy; device tensors
    ...                            # training_step(batch)
    ring.release(i)                # slot may be overwritten once the compute stream got here
    ring.wait_free(i)              # host: the copy out of host slot i is done, decode threads may refill it
"""
import torch


class TileRing:
    def __init__(self, batch, h, w, post=False, slots=2, device=None, extra=None):
        """`extra`: {name: (shape per sample, dtype)} of additional per-sample host arrays that travel with the tiles (e.g. the
        augmentation decisions of the device-side train augmentation)."""
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slots = slots
        self.post = post
        names = ["tiles", "mask"] + (["tiles_post"] if post else [])
        shape = {"tiles": (batch, h, w, 3), "tiles_post": (batch, h, w, 3), "mask": (batch, h, w)}
        dtypes = {k: torch.uint8 for k in names}
        for k, (shp, dt) in (extra or {}).items():
            names.append(k)
            shape[k] = (batch, *shp)
            dtypes[k] = dt
        self._host = [{k: torch.empty(shape[k], dtype=dtypes[k]).pin_memory() for k in names} for _ in range(slots)]
        self._dev = [{k: torch.empty(shape[k], dtype=dtypes[k], device=self.device) for k in names} for _ in range(slots)]
        self.stream = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event() for _ in range(slots)]
        self._free = [torch.cuda.Event() for _ in range(slots)]
        for e in self._free:
            e.record(torch.cuda.current_stream(self.device))
        self._submitted = [False] * slots
        self.batch, self.hw = batch, (h, w)
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in self._host[0].values())

    def host(self, i):
        """Pinned host tensors of slot i (fill them in place; `.numpy()` views share the memory)."""
        return self._host[i % self.slots]

    def wait_free(self, i):
        """Host-side: blocks until the last H2D copy out of host slot i has finished, so the slot may be refilled."""
        s = i % self.slots
        if self._submitted[s]:
            self._ready[s].synchronize()

    def submit(self, i):
        s = i % self.slots
        self._submitted[s] = True
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._free[s])
            for k, t in self._host[s].items():
                self._dev[s][k].copy_(t, non_blocking=True)
            self._ready[s].record(self.stream)

    def acquire(self, i):
        s = i % self.slots
        torch.cuda.current_stream(self.device).wait_event(self._ready[s])
        return self._dev[s]

    def release(self, i):
        self._free[i % self.slots].record(torch.cuda.current_stream(self.device))
