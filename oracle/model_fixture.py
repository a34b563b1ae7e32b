"""TEST INFRASTRUCTURE (not product code): a deterministic recipe for filling a backbone state_dict, shared by
oracle/make_model_golden.py (which fills the REFERENCE's MSPN2 with it) and tests/test_model.py (which rebuilds the
same tensors from the committed key/shape list and loads them through das_b200.model's checkpoint key map).
Storing the recipe instead of the tensors keeps the fixture small (the test backbone has ~20 M parameters)."""
import zlib

import torch


def synthetic_tensor(key: str, shape) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    shape = tuple(int(s) for s in shape)
    if key.endswith("running_var"):
        return 0.5 + torch.rand(shape, generator=g)
    if key.endswith("running_mean"):
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.long)
    if len(shape) == 4:                                   # convolution: keep activations O(1) through ~40 layers
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (1.5 / fan_in) ** 0.5
    if key.endswith("weight"):                            # norm scale
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.1 * torch.randn(shape, generator=g)          # biases


def synthetic_state(keys, shapes):
    return {k: synthetic_tensor(k, s) for k, s in zip(keys, shapes)}


def synthetic_image(batch: int, h: int, w: int, seed: int = 77) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, h, w, generator=g)
