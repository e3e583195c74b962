"""Host/device adaptation shared by the op mirrors: accept torch CUDA tensors (zero-copy), torch CPU
tensors or numpy arrays (copied to the device); give results back in the caller's flavour."""
import numpy as np
import torch

from ... import _lib


def dev(x, dtype=torch.float32):
    """-> (cuda tensor, flavour) ; flavour in {'cuda','cpu','numpy'}"""
    _lib.require_cuda()
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x.to(dtype).contiguous(), "cuda"
        return x.to(device="cuda", dtype=dtype, non_blocking=True).contiguous(), "cpu"
    a = np.ascontiguousarray(x)
    return torch.from_numpy(a).to("cuda", non_blocking=True).to(dtype).contiguous(), "numpy"


def back(t, flavour):
    if flavour == "cuda" or t is None:
        return t
    t = t.cpu()
    return t if flavour == "cpu" else t.numpy()
