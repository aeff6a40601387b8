"""Small host helpers: pointer/stream extraction, workspace allocation, argument
checks.  torch is used for device memory and streams only."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def ptr(t) -> int:
    """Raw device pointer of a tensor argument.  Every kernel is launched on the CURRENT device's current stream, so a
    tensor that lives on another GPU is rejected here instead of being dereferenced from the wrong device."""
    if t is None:
        return 0
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise ValueError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                         "select it first (torch.cuda.set_device / `with torch.cuda.device(...)`)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def need_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this library has no CPU path)")
    return t


def device_of(device=None) -> torch.device:
    if device is None:
        device = "cuda"
    d = torch.device(device)
    if d.type != "cuda":
        raise ValueError(f"device {d} requested: the distillation core runs on CUDA only")
    if d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    elif d.index != torch.cuda.current_device():
        raise ValueError(f"device {d} requested but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                         "select it first (torch.cuda.set_device / `with torch.cuda.device(...)`)")
    return d


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def pad4(n: int) -> int:
    return (int(n) + 3) // 4 * 4


def padded_rows(x: torch.Tensor) -> torch.Tensor:
    """Row-major f32 matrix whose leading dimension is a multiple of 4 and whose base is
    16-byte aligned (what the 128-bit gathers need).  Returns x itself when it already
    qualifies, else a zero-padded copy (view of width x.shape[1] into the padded buffer)."""
    assert x.dim() == 2 and x.dtype == torch.float32
    n, f = x.shape
    if x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.stride(0) >= pad4(f) and x.data_ptr() % 16 == 0:
        return x
    buf = torch.zeros((n, pad4(f)), dtype=torch.float32, device=x.device)
    buf[:, :f] = x
    return buf[:, :f]


def new_padded(n: int, f: int, device, zero: bool = False) -> torch.Tensor:
    alloc = torch.zeros if zero else torch.empty
    buf = alloc((n, pad4(f)), dtype=torch.float32, device=device)
    if not zero and pad4(f) != f:
        buf[:, f:] = 0
    return buf[:, :f]


def as_i32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.int32 else t.to(torch.int32)


def to_device_i64(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(device)


def to_device_f32(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def launch_count() -> int:
    return _lib.query("gdr_launch_count")
