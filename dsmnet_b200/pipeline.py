"""Host-buffer entry point of the inference paths: a double-buffered host -> device -> host pipeline around one CUDA-graph
replay per step (the end-to-end path is part of the package, not of the benchmark).

    pipe = HostPipeline(step_fn, [example_input, ...])        # step_fn(*device_inputs) -> sequence of device tensors
    t0 = pipe.submit(host_a, host_b)                          # pinned host tensors; returns a ticket
    t1 = pipe.submit(host_a2, host_b2)                        # H2D of step 1 overlaps the compute of step 0
    outs0 = pipe.wait(t0)                                     # pinned host tensors holding step 0's results

Three streams: copy-in, compute, copy-out.  Every step copies that step's inputs from (pinned) host memory into one of two
staging sets, copies them into the graph's static inputs, replays the graph, copies the results into one of two staging
sets and from there into one of two pinned host result sets.  A ticket's results stay valid until two further steps have
been submitted.  With `use_graph=False` (or when capture fails) the step function is launched eagerly — same kernels.
"""
from __future__ import annotations

import sys
from typing import Callable, List, Sequence

import torch


class HostPipeline:
    def __init__(self, step_fn: Callable, example_inputs: Sequence[torch.Tensor], use_graph: bool = True, warmup: int = 2):
        self.device = example_inputs[0].device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline runs CUDA inference paths only")
        self.step_fn = step_fn
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        cur = torch.cuda.current_stream(self.device)
        self.s_cmp.wait_stream(cur)
        self.graph = None
        with torch.cuda.stream(self.s_cmp), torch.no_grad():
            for _ in range(max(1, warmup)):
                outs = list(self.step_fn(*self.static_in))
            self.s_cmp.synchronize()
            if use_graph:
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.s_cmp, capture_error_mode="thread_local"):
                        outs = list(self.step_fn(*self.static_in))
                    self.graph = g
                except Exception as e:                                   # eager launches are still the same kernels
                    sys.stderr.write("HostPipeline: CUDA graph capture unavailable (%s); launching eagerly\n" % e)
                    self.graph = None
                    outs = list(self.step_fn(*self.static_in))
        self.static_out = outs
        self.stage_in = [[torch.empty_like(t) for t in self.static_in] for _ in range(2)]
        self.stage_out = [[torch.empty_like(t) for t in outs] for _ in range(2)]
        self.host_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs] for _ in range(2)]
        mk = lambda: [torch.cuda.Event() for _ in range(2)]
        self.ev_in, self.ev_used, self.ev_out, self.ev_done = mk(), mk(), mk(), mk()
        self.count = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.static_in)
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in outs)
        torch.cuda.synchronize(self.device)

    def submit(self, *host_inputs: torch.Tensor) -> int:
        i = self.count
        k = i & 1
        self.count += 1
        with torch.no_grad():
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_used[k])                    # step i-2 has consumed this staging set
                for dst, src in zip(self.stage_in[k], host_inputs):
                    dst.copy_(src, non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self.ev_in[k])
                for dst, src in zip(self.static_in, self.stage_in[k]):
                    dst.copy_(src, non_blocking=True)
                self.ev_used[k].record(self.s_cmp)
                if self.graph is not None:
                    self.graph.replay()
                    outs = self.static_out
                else:
                    outs = list(self.step_fn(*self.static_in))
                self.s_cmp.wait_event(self.ev_done[k])                   # step i-2's results have left this staging set
                for dst, src in zip(self.stage_out[k], outs):
                    dst.copy_(src, non_blocking=True)
                self.ev_out[k].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_out[k])
                for dst, src in zip(self.host_out[k], self.stage_out[k]):
                    dst.copy_(src, non_blocking=True)
                self.ev_done[k].record(self.s_out)
        return i

    def wait(self, ticket: int) -> List[torch.Tensor]:
        if ticket < self.count - 2 or ticket >= self.count:
            raise RuntimeError("HostPipeline: ticket %d is no longer (or not yet) available" % ticket)
        self.ev_done[ticket & 1].synchronize()
        return self.host_out[ticket & 1]

    def run(self, *host_inputs: torch.Tensor) -> List[torch.Tensor]:
        """one synchronous step: submit + wait"""
        return self.wait(self.submit(*host_inputs))
