"""Helpers for the -m gpu parity tests (device buffers via torch)."""
import functools

import numpy as np
import torch

from tests.common import PARAMS


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_host(t):
    return t.detach().cpu().numpy().view(np.uint64)


@functools.lru_cache(maxsize=4)
def gpu_ctx(name):
    from heongpu_b200 import api
    log_n, qb, pb = PARAMS[name]
    return api.HEContext(log_n, qb, pb, device=0)


@functools.lru_cache(maxsize=2)
def ref_gpu(name):
    """The reference's own CUDA kernels driven by oracle/ref_gpu_harness.cu."""
    from oracle import oracle as O, ref as R
    log_n, qb, pb = PARAMS[name]
    primes = O.generate_primes(1 << log_n, qb + pb)
    t = R.tables_for_refgpu(log_n, primes, len(qb), len(pb))
    return R.RefGpu(log_n, primes, len(qb), len(pb), t)
