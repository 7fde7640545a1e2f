"""CUDA-graph capture of the training step's device work.

One training step is 500-3 800 kernel launches issued from Python (ctypes + autograd bookkeeping, ~35 us each): at small
per-GPU batches (BASELINE config 4: 2 pairs) the launch rate, not the kernels, bounds the step.  ``GraphedTrainStep`` captures

    zero_grad -> Model.training_step (tile normalise, U-Net forward, loss) -> backward

    -> bucketed NCCL all-reduce of the flat gradient buffer (launched per network stage from backward hooks, on NCCL's own stream)
    -> fused optimizer (per-step scalars read from a device block) -> batched weight re-pack

once into a ``torch.cuda.CUDAGraph`` (all libxv2 launches go to torch's current stream, so they are recorded like any other
kernel; TMA descriptors are encoded on the host at capture time and stay valid because the graph's private memory pool
keeps every activation at a fixed address) and replays it per step.  Per step the host only refreshes the optimizer's 32-byte
scalar block (one async copy) and launches the graph.  ``full=False`` (or XV2_GRAPH_OPT=0) keeps the all-reduce, the optimizer
and the re-pack eager as in round 1.

    step = GraphedTrainStep(model, optimizer, batch)     # batch: dict of device tensors with the shapes of every later batch
    loss = step(batch)                                   # device tensor, valid until the next call

Stream rule (autograd): the gradient accumulators of the parameters belong to the stream on which the model FIRST ran.  If
eager steps precede the capture they must have run on a non-default stream and that stream must be passed as ``stream=``
(``Trainer.fit`` does this); capturing on a stream other than theirs would make the legacy stream depend on the capture.
"""
import torch

from . import lib


class GraphedTrainStep:
    def __init__(self, model, optimizer, batch, warmup=3, stream=None, full=None):
        import os
        self.model, self.optimizer, self.flat = model, optimizer, model.flat
        self.full = (os.environ.get("XV2_GRAPH_OPT", "1") != "0") if full is None else bool(full)
        if self.flat is None:
            raise lib.Xv2Error("configure_optimizers() must run before the step is captured (flat parameter buffer)")
        self.static = {k: torch.empty_like(v).copy_(v) for k, v in batch.items()}
        cur = torch.cuda.current_stream()
        self.stream = stream if stream is not None else torch.cuda.Stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):  # warm-up on the capture stream: allocator, packed weights, accumulators settle
            for _ in range(warmup):
                self._forward_backward()
                self._finish()
        cur.wait_stream(self.stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = lib.launches()
        if self.full:
            with torch.cuda.stream(self.stream):
                self.optimizer.prepare_step()  # the captured optimizer launch reads the device scalar block
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.loss = self._forward_backward()
            if self.full:
                self._finish_captured()
        if self.full:
            self.optimizer.step_count -= 1      # the capture itself does not execute: undo its bookkeeping
        self.launches_per_replay = lib.launches() - before

    def describe(self):
        if self.full:
            return (f"zero-grad + forward + backward + bucketed gradient all-reduce (overlapped with backward) + fused optimizer + "
                    f"weight re-pack replayed from ONE CUDA graph ({self.launches_per_replay} libxv2 launches per replay)")
        return (f"forward + backward replayed from one CUDA graph ({self.launches_per_replay} launches per replay); gradient "
                "all-reduce, fused AdamW and weight re-pack eager")

    def _finish_captured(self):
        n = self.flat.all_reduce_grads()
        self.optimizer.grad_scale = 1.0 / n
        self.optimizer.step_captured()

    def _forward_backward(self):
        from . import ops
        self.optimizer.zero_grad()
        with ops.defer_nbt():  # the 81 num_batches_tracked += 1 launches of the forward become one multi-tensor add
            loss = self.model.training_step(self.static, 0)
        loss.backward()
        ops.check_pending_addends()  # every gradient part parked by ops.fork has been collected by its BatchNorm backward
        return loss.detach()

    def _finish(self):
        n = self.flat.all_reduce_grads()
        self.optimizer.grad_scale = 1.0 / n
        self.optimizer.step()

    def __call__(self, batch):
        for k, dst in self.static.items():
            src = batch[k]
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self.full:
            self.optimizer.prepare_step()
            self.graph.replay()
            lib.add_launches(self.launches_per_replay)
            return self.loss
        self.graph.replay()
        lib.add_launches(self.launches_per_replay)
        self._finish()
        return self.loss
