"""Shared helpers for the parity tests: seeded synthetic inputs and norm-wise errors."""
import numpy as np
import torch


def synth_batch(B, O, A, seed=0, counts=False, dtype=torch.float32, p_term=0.05):
    """SURVEY.md section 8c determinism recipe: private generator, randn obs/next_obs,
    U(-1,1) actions, randn rewards, Bernoulli terminals."""
    g = torch.Generator().manual_seed(seed)
    batch = dict(
        observations=torch.randn(B, O, generator=g),
        next_observations=torch.randn(B, O, generator=g),
        actions=torch.rand(B, A, generator=g) * 2 - 1,
        rewards=torch.randn(B, 1, generator=g),
        terminals=(torch.rand(B, 1, generator=g) < p_term).float(),
    )
    if counts:
        batch['counts'] = torch.randint(0, 3, (B, 1), generator=g).float()
    return {k: v.to(dtype) for k, v in batch.items()}


def synth_eps(n, B, A, seed=1, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, A, generator=g).to(dtype) for _ in range(n)]


def rel_err(x, ref):
    """Norm-wise relative error ||x-ref|| / ||ref|| (SURVEY.md section 8d tolerances)."""
    x = torch.as_tensor(np.asarray(x), dtype=torch.float64).flatten()
    ref = torch.as_tensor(np.asarray(ref), dtype=torch.float64).flatten()
    d = torch.linalg.norm(x - ref)
    n = torch.linalg.norm(ref)
    return float(d / n) if n > 0 else float(d)


def max_abs(x, ref):
    x = torch.as_tensor(np.asarray(x), dtype=torch.float64)
    ref = torch.as_tensor(np.asarray(ref), dtype=torch.float64)
    return float((x - ref).abs().max())
