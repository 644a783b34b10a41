"""Per-module cache of engine plans keyed by input shapes and parameter versions."""
import torch


class _ParamWatch(object):
    """Remembers a module's parameter/buffer tensors and notices in-place modification (autograd version
    counters) or re-homing (.to()/.cuda(): storage pointers of the end-point tensors) without walking the
    module tree on every call.  `invalidate_plans(module)` forces a refresh (tensors re-assigned by hand)."""

    def __init__(self, module):
        self.module, self.epoch = module, 0
        self.refresh()

    def refresh(self):
        self.ts = list(self.module.parameters()) + list(self.module.buffers())
        self.vers = [t._version for t in self.ts]
        self.ptrs = (self.ts[0].data_ptr(), self.ts[-1].data_ptr()) if self.ts else (0, 0)
        self.epoch += 1

    def stamp(self):
        ts, vers = self.ts, self.vers
        dirty = False
        for i in range(len(ts)):
            if ts[i]._version != vers[i]:
                dirty = True
                break
        if not dirty and ts and (ts[0].data_ptr(), ts[-1].data_ptr()) != self.ptrs:
            dirty = True
        if dirty:
            self.refresh()
        return self.epoch


def param_stamp(module):
    """Small integer that changes whenever the module's weights may have changed."""
    w = module.__dict__.get("_iou_param_watch")
    if w is None:
        w = _ParamWatch(module)
        module.__dict__["_iou_param_watch"] = w
    return w.stamp()


def invalidate_plans(module):
    """Call after replacing parameter tensors by hand (module.weight = nn.Parameter(...))."""
    for m in module.modules():
        w = m.__dict__.get("_iou_param_watch")
        if w is not None:
            w.refresh()


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
