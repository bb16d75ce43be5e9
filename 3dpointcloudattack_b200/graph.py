"""CUDA-graph capture of a distance-loss forward + backward (fixed shapes).

The attack loops call the same distance loss thousands of times with the same shapes
(attack/CW/CW_attack.py:111-178: 10 binary steps x 500 iterations).  In eager mode every call
pays Python / autograd / allocator overhead that is several times the GPU time of the kernels;
capturing forward + backward once and replaying the graph removes it.  All kernels of this
package are capture-safe: stream-ordered, allocation-free, no host synchronisation.
"""
import torch


class GraphedLoss:
    """Capture `loss, aux = fn(adv, ori)` and `loss.backward()` into one CUDA graph.

    fn(adv, ori) -> (scalar loss tensor, tuple of auxiliary tensors to keep)
    replay(adv=None, ori=None) copies new inputs into the static buffers (if given), replays the
    graph and returns (loss, aux, grad_adv) -- static tensors that the next replay overwrites.
    """

    def __init__(self, fn, adv, ori, warmup=3):
        if not adv.is_cuda:
            raise RuntimeError("GraphedLoss needs CUDA tensors (no CPU fallback)")
        self.fn = fn
        self.adv = adv.detach().clone().requires_grad_(True)
        self.ori = ori.detach().clone()
        # d(loss)/d(loss) = 1 as a static tensor: loss.backward() would launch a fill kernel on every replay
        self._one = torch.ones((), dtype=torch.float32, device=adv.device)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.adv.grad = None
                loss, _ = fn(self.adv, self.ori)
                torch.autograd.backward(loss, grad_tensors=self._one.expand_as(loss))
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.adv.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):      # same stream as the warm-up (AccumulateGrad node)
            self.loss, self.aux = fn(self.adv, self.ori)
            torch.autograd.backward(self.loss, grad_tensors=self._one.expand_as(self.loss))
        self.grad = self.adv.grad

    def replay(self, adv=None, ori=None):
        if adv is not None:
            with torch.no_grad():
                self.adv.copy_(adv, non_blocking=True)
        if ori is not None:
            self.ori.copy_(ori, non_blocking=True)
        self.graph.replay()
        return self.loss, self.aux, self.grad


def tapered_edges(B, chunks):
    """Slice boundaries of PipelinedLoss: the first slice is two thirds the size of the others (weights 2:3:3:...) --
    nothing can run until the first upload has landed, so a short first slice starts the kernels earlier (B=32, three
    slices: 8 + 12 + 12; measured 219 us per step vs 226 us for 10 + 11 + 11)."""
    chunks = max(1, min(int(chunks), B))
    w = [2] + [3] * (chunks - 1)
    tot, acc, edges = sum(w), 0, [0]
    for c in range(chunks):
        acc += w[c]
        e = B if c == chunks - 1 else int(round(B * acc / tot))
        e = min(max(e, edges[-1] + 1), B - (chunks - 1 - c))       # every slice keeps at least one sample
        edges.append(e)
    return edges


class PipelinedLoss:
    """Host-to-host step of a per-sample-separable distance loss as ONE CUDA graph with two (or more)
    internal branches: the batch is cut into `chunks` slices, every slice has its own stream inside the
    capture -- host->device copy from the pinned input buffers, forward + backward, device->host copy
    into the pinned output buffers -- so the copies of one slice overlap the kernels of the other
    (copy engines and SMs are independent; a monolithic step leaves the SMs idle during 3 MB in +
    1.5 MB out).  Driving the slices from Python instead is host-launch bound and slower than the
    monolithic step; inside one graph the branches cost one launch.

    fn(adv, ori) -> (scalar loss, aux tuple): must treat the samples independently (every loss of this
    package does).
    Buffers (static, pinned): adv_host, ori_host [B, ...] inputs (write new data into them in place),
    grad_host [B, ...], aux_host[c][i] = the i-th aux tensor of slice c (contiguous per slice: a strided
    device->host copy is not capturable; `slices` lists the sample ranges).  replay() runs the graph; the
    results are valid after a synchronisation of the current stream.
    """

    def __init__(self, fn, adv_host, ori_host, chunks=2, warmup=3, device=None, slice_sizes=None, prioritize=True):
        if not (adv_host.is_pinned() and ori_host.is_pinned()):
            raise ValueError("PipelinedLoss needs pinned host tensors")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        B = adv_host.shape[0]
        if slice_sizes is not None:
            if sum(slice_sizes) != B or min(slice_sizes) <= 0:
                raise ValueError("slice_sizes must be positive and sum to the batch size")
            edges = [0]
            for s in slice_sizes:
                edges.append(edges[-1] + int(s))
        else:
            edges = tapered_edges(B, chunks)
        self.slices = list(zip(edges[:-1], edges[1:]))
        self.adv_host, self.ori_host = adv_host, ori_host
        self.grad_host = torch.empty_like(adv_host).pin_memory()
        self.adv_dev = [adv_host[lo:hi].to(dev).requires_grad_(True) for lo, hi in self.slices]
        self.ori_dev = [ori_host[lo:hi].to(dev) for lo, hi in self.slices]
        cur = torch.cuda.current_stream(dev)
        main = torch.cuda.Stream(dev)
        # earlier slices get the higher stream priority: when SM slots free up, the tail of slice c (fix-up,
        # reductions, backward -- which release its device->host copies) is placed before the sweep of c+1
        try:
            lo_pri, hi_pri = torch.cuda.Stream.priority_range()      # (least, greatest), e.g. (0, -5)
        except Exception:
            lo_pri, hi_pri = 0, -1
        n_sl = len(self.slices)
        sides = [torch.cuda.Stream(dev, priority=(max(hi_pri, lo_pri - (n_sl - 1 - c)) if prioritize else 0))
                 for c in range(n_sl)]
        copiers = [torch.cuda.Stream(dev, priority=st.priority) for st in sides]
        aux_shapes = [None] * len(sides)
        for c, st in enumerate(sides):                    # warm-up on the stream the slice will be captured on
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for _ in range(warmup):
                    self.adv_dev[c].grad = None
                    loss, aux = fn(self.adv_dev[c], self.ori_dev[c])
                    loss.backward()
                aux_shapes[c] = [(tuple(a.shape), a.dtype) for a in aux]
            cur.wait_stream(st)
        torch.cuda.synchronize(dev)
        self.aux_host = [[torch.empty(shape, dtype=dtype).pin_memory() for shape, dtype in per] for per in aux_shapes]
        self._one = torch.ones((), dtype=torch.float32, device=dev)
        for a in self.adv_dev:
            a.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=main):
            prev_h2d = None
            for c, ((lo, hi), st) in enumerate(zip(self.slices, sides)):
                st.wait_stream(main)                                       # fork
                if prev_h2d is not None:
                    st.wait_event(prev_h2d)                                # uploads strictly in slice order
                with torch.cuda.stream(st):
                    with torch.no_grad():
                        self.adv_dev[c].copy_(adv_host[lo:hi], non_blocking=True)
                        self.ori_dev[c].copy_(ori_host[lo:hi], non_blocking=True)
                    prev_h2d = torch.cuda.Event()
                    prev_h2d.record(st)
                    loss, aux = fn(self.adv_dev[c], self.ori_dev[c])
                    aux = tuple(a.detach() for a in aux)
                    cp = copiers[c]                                        # forward results leave on a branch of
                    cp.wait_stream(st)                                     # their own while the backward runs: only
                    with torch.cuda.stream(cp):                            # the gradient copy trails the kernels
                        for out, a in zip(self.aux_host[c], aux):
                            out.copy_(a, non_blocking=True)
                            a.record_stream(cp)
                    torch.autograd.backward(loss, grad_tensors=self._one.expand_as(loss))
                    self.grad_host[lo:hi].copy_(self.adv_dev[c].grad, non_blocking=True)
                    st.wait_stream(cp)
            for st in sides:
                main.wait_stream(st)                                       # join
        self._main = main

    def replay(self):
        self.graph.replay()
        return self.aux_host, self.grad_host
