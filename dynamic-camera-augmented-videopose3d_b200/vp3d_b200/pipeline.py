"""Host <-> device pipelining around the hot path (copy / compute overlap on separate CUDA streams).

The reference moves every batch with a blocking .cuda() and reads results back with .cpu() (run.py:458-464,698-705,
732), so PCIe time adds to compute time. Here copies run on their own streams:

* infer_host(): a batch of sequences held in pinned host memory is cut into chunks of sequences; chunk i+1 is uploaded
  and chunk i-1 is downloaded while chunk i runs through the model (three streams, event-ordered, two buffers each).
* HostInferPipeline: the same across batches -- nothing drains between two batches of a serving loop.
* HostPrefetcher: double-buffered upload of training batches -- the next batch travels while the current step runs.
"""
import torch


def infer_host(model, x_host, y_host=None, chunk_seqs=8):
    """x_host: pinned (N, T, J, F) fp32 -> y_host: pinned (N, T', J_out, 3) fp32; model in eval mode on a CUDA device.
    Returns y_host once all results have landed (the call synchronises at the end)."""
    assert not model.training, 'infer_host runs the eval-mode (folded BatchNorm) path'
    dev = next(model.parameters()).device
    n = x_host.shape[0]
    t_out = x_host.shape[1] - (model.receptive_field() - 1)
    if y_host is None:
        y_host = torch.empty((n, t_out, model.num_joints_out, 3), dtype=torch.float32).pin_memory()
    main = torch.cuda.current_stream(dev)
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    up.wait_stream(main)
    down.wait_stream(main)
    chunks = [(lo, min(lo + chunk_seqs, n)) for lo in range(0, n, chunk_seqs)]
    xbuf = [torch.empty((chunk_seqs,) + tuple(x_host.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
    x_free = [torch.cuda.Event() for _ in range(2)]      # compute has consumed buffer b
    x_ready = [torch.cuda.Event() for _ in range(2)]     # upload into buffer b finished
    results = []

    def upload(i):
        lo, hi = chunks[i]
        b = i % 2
        with torch.cuda.stream(up):
            if i >= 2:
                up.wait_event(x_free[b])
            xbuf[b][:hi - lo].copy_(x_host[lo:hi], non_blocking=True)
            x_ready[b].record(up)

    upload(0)
    with torch.no_grad():
        for i, (lo, hi) in enumerate(chunks):
            b = i % 2
            if i + 1 < len(chunks):
                upload(i + 1)
            main.wait_event(x_ready[b])
            y = model(xbuf[b][:hi - lo])
            x_free[b].record(main)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(down):
                down.wait_event(done)
                y_host[lo:hi].copy_(y, non_blocking=True)
            y.record_stream(down)
            results.append(y)
    down.synchronize()
    main.synchronize()
    return y_host


class HostInferPipeline:
    """infer_host without the drain at the end of every batch: streams, device staging buffers and the chunk counter
    persist between calls, so the first upload of batch i+1 and the last download of batch i overlap with compute of
    the neighbouring batch. `submit(x_host, y_host)` enqueues one batch and returns a CUDA event that fires when its
    last result byte is in `y_host` (pinned); `infer()` = submit + wait. Successive calls must pass distinct `y_host`
    buffers for as long as an earlier result is still being read."""

    def __init__(self, model, chunk_seqs=8):
        assert not model.training, 'HostInferPipeline runs the eval-mode (folded BatchNorm) path'
        self.model = model
        self.dev = next(model.parameters()).device
        self.chunk_seqs = chunk_seqs
        self.up, self.down = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.xbuf = None
        self.x_free = [torch.cuda.Event() for _ in range(2)]
        self.x_ready = [torch.cuda.Event() for _ in range(2)]
        self.n_chunks = 0        # chunks issued so far, over all batches (buffer = chunk number % 2)

    def submit(self, x_host, y_host):
        model, dev, cs = self.model, self.dev, self.chunk_seqs
        n = x_host.shape[0]
        main = torch.cuda.current_stream(dev)
        if self.xbuf is None or tuple(self.xbuf[0].shape[1:]) != tuple(x_host.shape[1:]):
            torch.cuda.synchronize(dev)
            self.xbuf = [torch.empty((cs,) + tuple(x_host.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
            self.n_chunks = 0
        chunks = [(lo, min(lo + cs, n)) for lo in range(0, n, cs)]

        def upload(k, lo, hi):
            b = k % 2
            with torch.cuda.stream(self.up):
                if k >= 2:
                    self.up.wait_event(self.x_free[b])
                self.xbuf[b][:hi - lo].copy_(x_host[lo:hi], non_blocking=True)
                self.x_ready[b].record(self.up)

        k0 = self.n_chunks
        upload(k0, *chunks[0])
        with torch.no_grad():
            for i, (lo, hi) in enumerate(chunks):
                k = k0 + i
                b = k % 2
                if i + 1 < len(chunks):
                    upload(k + 1, *chunks[i + 1])
                main.wait_event(self.x_ready[b])
                y = model(self.xbuf[b][:hi - lo])
                self.x_free[b].record(main)
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(self.down):
                    self.down.wait_event(done)
                    y_host[lo:hi].copy_(y, non_blocking=True)
                y.record_stream(self.down)
        self.n_chunks = k0 + len(chunks)
        landed = torch.cuda.Event()
        landed.record(self.down)
        return landed

    def infer(self, x_host, y_host=None):
        if y_host is None:
            t_out = x_host.shape[1] - (self.model.receptive_field() - 1)
            y_host = torch.empty((x_host.shape[0], t_out, self.model.num_joints_out, 3), dtype=torch.float32).pin_memory()
        self.submit(x_host, y_host).synchronize()
        return y_host


class HostPrefetcher:
    """Double-buffered host -> device upload. `put(tensors)` starts the copy of a batch of pinned host tensors on a side
    stream; `get()` makes the current stream wait for it and returns the device tensors (valid until the next get())."""

    def __init__(self, example_tensors, device):
        self.dev = device
        self.stream = torch.cuda.Stream(device)
        self.bufs = [[torch.empty_like(t, device=device) for t in example_tensors] for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.put_idx = 0
        self.get_idx = 0
        self.used = [False, False]

    def put(self, host_tensors):
        b = self.put_idx % 2
        with torch.cuda.stream(self.stream):
            if self.used[b]:
                self.stream.wait_event(self.free[b])
            for dst, src in zip(self.bufs[b], host_tensors):
                dst.copy_(src, non_blocking=True)
            self.ready[b].record(self.stream)
        self.put_idx += 1

    def get(self):
        b = self.get_idx % 2
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ready[b])
        self.get_idx += 1
        return b, self.bufs[b]

    def release(self, b):
        """The current stream has finished reading buffer b (call after the consumer work has been enqueued)."""
        self.free[b].record(torch.cuda.current_stream(self.dev))
        self.used[b] = True
