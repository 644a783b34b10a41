"""Per-module cache of engine plans keyed by input shapes and parameter versions."""
import torch


def param_stamp(module):
    """Changes whenever a parameter/buffer is modified in place or replaced."""
    s = 0
    for t in list(module.parameters()) + list(module.buffers()):
        s = (s * 1000003 + t._version + (t.data_ptr() % 1000003)) % (1 << 61)
    return s


def cuda_state_dict(module, device):
    return {k: v.detach().to(device) for k, v in module.state_dict().items()}


class PlanCache(object):
    def __init__(self, max_plans=2):
        self.plans = {}
        self.max_plans = max_plans

    def get(self, key, builder):
        if key not in self.plans:
            if len(self.plans) >= self.max_plans:
                self.plans.pop(next(iter(self.plans)))
            self.plans[key] = builder()
        return self.plans[key]

    def clear(self):
        self.plans.clear()


def require_cuda(x, what):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("%s: expected a CUDA tensor -- this path has no CPU fallback" % what)
