"""Drop-in for ``everyvoice/preprocessor/helpers.py:47-106`` (``Scaler``) on a B200.

``append`` keeps tensors (device or host); ``calculate_stats`` reduces each of them on the
GPU to ``{count, sum, sumsq, min, max}`` (NaN-skipping, float64 accumulation,
``evf_stats_partial``), optionally all-reduces those five numbers across the ranks of a
process group, and derives the reference's statistics.  ``normalize`` is the reference's
out-of-place ``(x - mean) / std``; ``normalize_`` does a whole shard in place.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .distributed import allreduce_stats, finalize_stats
from .heavy import _ptr, _require_cuda, _stream_ptr


class Scaler:
    def __init__(self, device=None):
        self._device = device
        self._data: list[torch.Tensor] = []
        self._stats5 = None
        self.min = None
        self.max = None
        self.std = None
        self.mean = None
        self.norm_min = None
        self.norm_max = None

    def __len__(self):
        return len(self.data)

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        raise ValueError(
            f"Sorry, you tried to change the data to {value} but it cannot be changed directly. Either Scaler.append(data), or Scaler.clear_data()"
        )

    def append(self, value):
        self._data.append(value)
        self._stats5 = None

    def clear_data(self):
        """Clear data"""
        self.__init__(self._device)

    # -- device reductions -------------------------------------------------------------------
    def _device_for(self, t: torch.Tensor | None = None) -> torch.device:
        if t is not None and t.is_cuda:
            return t.device
        return _require_cuda(self._device)

    def partial_stats(self) -> torch.Tensor:
        """float64[5] ``{count, sum, sumsq, min, max}`` over everything appended, on the GPU."""
        if self._stats5 is None:
            lib = _lib.load()
            device = self._device_for(self._data[0] if self._data else None)
            out = torch.empty(5, dtype=torch.float64, device=device)
            with torch.cuda.device(device):
                first = True
                for t in self._data:
                    td = t.to(device=device, dtype=torch.float32).contiguous().view(-1)
                    _lib.check(lib.evf_stats_partial(_ptr(td), td.numel(), _ptr(out), 0 if first else 1, _stream_ptr(device)))
                    first = False
                if first:
                    _lib.check(lib.evf_stats_partial(C.c_void_p(0), 0, _ptr(out), 0, _stream_ptr(device)))
            self._stats5 = out
        return self._stats5

    def calculate_stats(self, group=None, distributed: bool | None = None):
        """Reference: helpers.py:86-106.  With ``distributed`` (default: whenever
        ``torch.distributed`` is initialised) the five partial sums and the file count are
        all-reduced over ``group`` first, so every rank returns the corpus-wide statistics."""
        import torch.distributed as dist

        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized()
        if not len(self) and not distributed:
            return
        stats5, sample_size = self.partial_stats(), len(self)
        if distributed:
            stats5, sample_size = allreduce_stats(stats5, sample_size, group)
        stats = finalize_stats(stats5.cpu().tolist(), sample_size)
        device = stats5.device
        for k in ("min", "max", "mean", "std", "norm_min", "norm_max"):
            setattr(self, k, torch.tensor(stats[k], dtype=torch.float32, device=device))
        return stats

    def normalize(self, data):
        """Remove mean and normalize to unit variance"""
        if not torch.is_tensor(data):
            return (data - float(self.mean)) / float(self.std)
        device = self._device_for(data)
        out = data.to(device=device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        self.normalize_(out)
        return out if data.is_cuda else out.to(data.device)

    def normalize_(self, data: torch.Tensor) -> torch.Tensor:
        """In place over a contiguous float32 device tensor (a whole shard in one launch)."""
        if not (data.is_cuda and data.dtype == torch.float32 and data.is_contiguous()):
            raise ValueError("normalize_ needs a contiguous float32 CUDA tensor")
        lib = _lib.load()
        with torch.cuda.device(data.device):
            _lib.check(lib.evf_normalize_inplace(_ptr(data), data.numel(), float(self.mean), float(self.std),
                                                 _stream_ptr(data.device)))
        return data

    def normalize_by_device_stats_(self, data: torch.Tensor, stats5: torch.Tensor) -> torch.Tensor:
        """In place, with mean / std derived on the device from the float64 summary -- either one
        ``[5]`` summary or the ``[world, 5]`` output of ``distributed.allgather_stats``: the
        reduction, the collective and the normalisation stay on the stream."""
        if not (data.is_cuda and data.dtype == torch.float32 and data.is_contiguous()):
            raise ValueError("normalize_by_device_stats_ needs a contiguous float32 CUDA tensor")
        if not (stats5.is_cuda and stats5.dtype == torch.float64 and stats5.is_contiguous() and stats5.numel() >= 5):
            raise ValueError("stats5 must be a contiguous float64 CUDA tensor of [5] or [world, 5]")
        lib = _lib.load()
        with torch.cuda.device(data.device):
            if stats5.dim() == 2:
                _lib.check(lib.evf_normalize_by_gathered_stats(_ptr(data), data.numel(), _ptr(stats5), stats5.shape[0],
                                                               stats5.shape[1], _stream_ptr(data.device)))
            else:
                _lib.check(lib.evf_normalize_by_stats(_ptr(data), data.numel(), _ptr(stats5), _stream_ptr(data.device)))
        return data

    def denormalize(self, data):
        """Get de-normalized value"""
        return (data * self.std.to(data.device)) + self.mean.to(data.device) if torch.is_tensor(data) \
            else (data * float(self.std)) + float(self.mean)
