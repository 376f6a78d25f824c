"""Concurrent synthesis of independent batches on several CUDA streams of one GPU.

The reference's driver loops over a LIST of batches, one forward at a time (`synthesize.py:59-76`:
`for batch in batchs: ... output = model(*(batch[2:]))`).  The batches are independent, so this module runs them on
`n_streams` streams at once: each stream has its own worker thread and its own engine (C handle: packed weights and
workspace; `FastSpeech2Align` keeps one per stream), and a forward's host-side work -- the read-back of T between the
two stages, output allocation, the H2D / D2H copies of host batches -- overlaps the other streams' kernels.  On the
GPU, kernels of different forwards fill each other's tails: a persistent GEMM whose last round occupies 6 of 148 SMs
no longer idles the rest (at batch 32 the decoder has 151 row tiles).  Results are bit-identical to sequential calls
(tests/test_gpu_streamed.py).  Measured on B200 (bench.py): batch 32, 1 -> 3 streams: 11.0 -> 15.9 M frames/s with
device-resident inputs, 9.9 -> 15.0 M frames/s end to end from pinned host buffers.  Callers should consume and drop
results as they arrive: every live result tuple pins ~20 MB of device (and, with to_host, page-locked host) memory that
later forwards would otherwise reuse from the allocator caches.
"""
from __future__ import annotations

import queue
import threading
from typing import List, Optional, Sequence, Tuple

import torch

Batch = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]   # speakers[B], texts[B,L], src_lens[B], max_src_len


class _Job:
    __slots__ = ("batch", "kw", "to_host", "post", "result", "error", "done", "start_event")

    def __init__(self, batch, kw, to_host, start_event, post=None):
        self.batch, self.kw, self.to_host, self.start_event, self.post = batch, kw, to_host, start_event, post
        self.result, self.error = None, None
        self.done = threading.Event()


class StreamedSynthesizer:
    """`run(batches)` = `[model(*b) for b in batches]`, executed on `n_streams` CUDA streams concurrently.

    Batches may hold host tensors (ideally pinned: the H2D copies are then asynchronous) or device tensors.  With
    `to_host=True` (every tensor) or `to_host=(1, 9)` (those positions of the 12-tuple; the rest stay on the device) the
    worker copies results into pinned host memory: the D2H is part of the job and overlaps other streams' compute.

    Device results are complete when `wait` returns (the worker synchronises its stream), but their memory belongs to
    that stream's allocator pool: a caller that consumes them asynchronously on ANOTHER stream and drops them before that
    work has run should call `tensor.record_stream(consumer_stream)` first (the usual torch multi-stream rule).
    `close()` frees the per-stream engines (packed weights + workspaces) this synthesizer created inside the module."""

    def __init__(self, model, n_streams: int = 3, device: Optional[torch.device] = None):
        if n_streams < 1:
            raise ValueError("n_streams must be >= 1")
        self.model = model
        self.device = device if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("StreamedSynthesizer needs the module on a CUDA device")
        self.n_streams = n_streams
        self._streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self._queue: "queue.Queue[Optional[_Job]]" = queue.Queue()
        # the workers hold the queue, their stream and the model -- NOT this object: a synthesizer that is dropped without
        # close() is still collected, and its __del__ stops the threads and releases the engines
        self._threads = [threading.Thread(target=self._worker, args=(self._queue, self._streams[i], model, self.device),
                                          daemon=True) for i in range(n_streams)]
        for t in self._threads:
            t.start()

    def __enter__(self) -> "StreamedSynthesizer":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # ------------------------------------------------------------------ worker
    @staticmethod
    def _worker(jobs: "queue.Queue[Optional[_Job]]", stream: torch.cuda.Stream, model, device: torch.device) -> None:
        torch.cuda.set_device(device)
        while True:
            job = jobs.get()
            if job is None:
                return
            try:
                with torch.cuda.stream(stream), torch.no_grad():
                    if job.start_event is not None:
                        stream.wait_event(job.start_event)
                    sp, tx, sl, L = job.batch
                    sp, tx, sl = (t.to(device, non_blocking=True) for t in (sp, tx, sl))
                    out, info = model.forward_with_info(sp, tx, sl, L, **job.kw)
                    if job.post is not None:      # hand-off work (pipeline.py) runs on this job's stream as well
                        out = job.post(out, info)
                    elif job.to_host:
                        sel = None if job.to_host is True else set(job.to_host)
                        host = []
                        for k, t in enumerate(out):
                            if t is None or (sel is not None and k not in sel):
                                host.append(t)
                                continue
                            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                            h.copy_(t, non_blocking=True)
                            host.append(h)
                        out = tuple(host)
                    stream.synchronize()          # the job is done when its results are usable by the caller
                    job.result = out
            except BaseException as e:            # surfaced by run() in the calling thread
                job.error = e
            finally:
                job.done.set()

    # ------------------------------------------------------------------ API
    def submit(self, batch: Batch, to_host=False, start_event: Optional[torch.cuda.Event] = None, post=None, **kw) -> _Job:
        """`post(predictions, info)`: optional callable run by the worker on the job's stream right after the forward
        (`info` = {"T", "frames"}, see FastSpeech2Align.forward_with_info); its return value becomes the job's result."""
        job = _Job(batch, kw, to_host, start_event, post)
        self._queue.put(job)
        return job

    @staticmethod
    def wait(job: _Job):
        job.done.wait()
        if job.error is not None:
            raise job.error
        return job.result

    def run(self, batches: Sequence[Batch], to_host=False, **kw) -> List[tuple]:
        """Results in the order of `batches`."""
        jobs = [self.submit(b, to_host=to_host, **kw) for b in batches]
        return [self.wait(j) for j in jobs]

    def warm_up(self, batch: Batch) -> None:
        """One forward per stream: creates the per-stream engines (weight repacking) and sizes their workspaces."""
        barrier = threading.Barrier(self.n_streams)
        errs = []

        def one(i):
            try:
                with torch.cuda.stream(self._streams[i]), torch.no_grad():
                    sp, tx, sl, L = batch
                    barrier.wait()
                    self.model(*(t.to(self.device) for t in (sp, tx, sl)), L)
                    self._streams[i].synchronize()
            except BaseException as e:
                errs.append(e)

        ts = [threading.Thread(target=one, args=(i,)) for i in range(self.n_streams)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]

    def close(self) -> None:
        """Stop the workers (after the jobs already queued) and free the engines (workspaces; the packed weights are
        shared with the module's other engines) of this synthesizer's streams.  Blocks until every worker has exited: an
        engine must not be released under a forward that is still running.  Idempotent."""
        if not self._threads:
            return
        for _ in self._threads:
            self._queue.put(None)
        for t in self._threads:
            t.join()
        self._threads = []
        if hasattr(self.model, "release_engine"):
            for st in self._streams:
                st.synchronize()
                self.model.release_engine(st)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
