"""Deterministic, NAME-keyed parameter initialisation shared by the golden generator (genuine
reference on the shims), the oracle and the CUDA product, so that all three hold identical
weights without storing them.  Values are fp32-representable (drawn in fp32, then cast)."""
import math
import zlib

import torch


def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def reseed_parameters(model, seed, strip_prefixes=("module.",)):
    with torch.no_grad():
        for name, p in model.named_parameters():
            key = name
            for pre in strip_prefixes:
                if key.startswith(pre):
                    key = key[len(pre):]
            g = _gen(key, seed)
            r = torch.randn(p.numel(), generator=g, dtype=torch.float32).reshape(p.shape)
            if key.endswith("bessel_weights"):
                n = p.numel()
                val = torch.linspace(1.0, n, n) * math.pi + 0.05 * r
            elif key.endswith("norm.std"):
                val = 1.0 + 0.1 * r
            elif key.endswith("bias"):
                val = 0.1 * r
            else:
                val = r
            p.copy_(val.to(dtype=p.dtype, device=p.device))
    return model
