"""CUDA-graph capture of a learner's device-side update.

The launch-bound learners (SAC / DQN / QMIX at the BASELINE batch sizes: ~100-300 kernel launches of a few
microseconds each per update) spend their time in Python and launch latency, not on the GPU.  ``CapturedStep`` records
the device half of an update once - network forward/backward (cuDNN / cuBLAS), the xb200 kernels (they launch on
torch's current stream, so capture sees them) and the flat-bucket Adam kernels - and replays it with one
``cudaGraphLaunch`` per update.  Host-side work (step counters, learning-rate schedule, the 16-byte hyper-parameter
upload, target sync) stays outside the graph."""
import torch


class CapturedStep:
    def __init__(self, fn, example_inputs, snapshot, restore, warmup=2):
        """fn(*tensors) -> tuple of tensors (or None); it must be pure device work with static shapes.
        snapshot() / restore(s) save and reinstate every piece of state ``fn`` mutates (warm-up runs are real)."""
        self.static_in = [x.clone() if isinstance(x, torch.Tensor) else x for x in example_inputs]
        snap = snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        restore(snap)
        torch.cuda.synchronize()
        from ... import _lib
        self._lib = _lib
        before = _lib.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)
        self.xb_launches = _lib.launch_count - before     # xb200 kernels inside one replay
        _lib.launch_count = before

    def __call__(self, *inputs):
        for s, x in zip(self.static_in, inputs):
            if isinstance(s, torch.Tensor) and s.data_ptr() != x.data_ptr():
                s.copy_(x, non_blocking=True)
        self.graph.replay()
        self._lib.launch_count += self.xb_launches
        return self.static_out
