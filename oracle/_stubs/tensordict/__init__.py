"""Minimal stand-in for ``tensordict.TensorDict`` (absent in this image) -- just enough for the
reference's Torch closures (src/torchrl/reppo.py:143-158,201-219,621-623) to run on CPU.
Test infrastructure only."""
import torch


class TensorDict(dict):
    def __init__(self, source=None, batch_size=None, device=None):
        super().__init__(source or {})
        self.batch_size = tuple(batch_size) if batch_size is not None else ()
        self.device = device

    def _map(self, fn, batch_size=None):
        return TensorDict({k: fn(v) for k, v in self.items()},
                          batch_size=self.batch_size if batch_size is None else batch_size,
                          device=self.device)

    def float(self):
        return self._map(lambda v: v.float())

    def detach(self):
        return self._map(lambda v: v.detach())

    def contiguous(self):
        return self._map(lambda v: v.contiguous())

    def flatten(self, a, b):
        bs = self.batch_size
        n = 1
        for d in bs[a:b + 1]:
            n *= d
        return self._map(lambda v: v.flatten(a, b), batch_size=bs[:a] + (n,) + bs[b + 1:])

    def __getitem__(self, key):
        if isinstance(key, str):
            return dict.__getitem__(self, key)
        out = self._map(lambda v: v[key])
        first = next(iter(out.values()), None)
        if first is not None:
            out.batch_size = tuple(first.shape[:1])
        return out
