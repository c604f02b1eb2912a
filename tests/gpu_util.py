"""Helpers for the -m gpu tests: numpy (oracle storage format) <-> torch CUDA tensors."""

import numpy as np
import torch

TORCH_DTYPE = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def to_torch(a: np.ndarray, dtype: str, device="cuda") -> torch.Tensor:
    if dtype == "bf16":
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).to(device)
    return torch.from_numpy(a.copy()).to(device)


def to_numpy(t: torch.Tensor, dtype: str) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if dtype == "bf16":
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def kv_to_torch(layers, dtype, device="cuda"):
    return [(to_torch(K, dtype, device), to_torch(V, dtype, device)) for K, V in layers]
