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
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.adv.grad = None
                loss, _ = fn(self.adv, self.ori)
                loss.backward()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.adv.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):      # same stream as the warm-up (AccumulateGrad node)
            self.loss, self.aux = fn(self.adv, self.ori)
            self.loss.backward()
        self.grad = self.adv.grad

    def replay(self, adv=None, ori=None):
        if adv is not None:
            with torch.no_grad():
                self.adv.copy_(adv, non_blocking=True)
        if ori is not None:
            self.ori.copy_(ori, non_blocking=True)
        self.graph.replay()
        return self.loss, self.aux, self.grad
