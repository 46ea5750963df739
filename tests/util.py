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


def per_sample_flip_split(got_w, ref_w, got_b, ref_b, X, thresh=1e-3):
    """Flip-aware comparison of a FIRST-layer weight gradient dW0 = dh1^T X  ([H, K], X = the layer's [B, K] input).

    Two correct evaluations of the same ReLU network whose pre-activations differ by a relative 1e-5 .. 1e-4 (the tcgen05
    accumulator truncates, so kernel and CPU model differ by that much) disagree on the ReLU mask of the one or two
    (sample, unit) pairs that sit that close to zero; one such flip in layer 2 changes dh1 of THAT SAMPLE in every unit,
    i.e. every row of dW0 -- ~5e-3 norm-wise -- although nothing is wrong.  With B < K the rows of X are independent,
    so the per-sample factors are recoverable: dW0 pinv(X) = dh1^T.  Returns (n_bad, clean_w, clean_b): the number of
    samples whose dh1 differs by more than ``thresh`` (relative), and the norm-wise errors of the weight / bias gradient
    with those samples' contributions taken out of both sides."""
    Xd = torch.as_tensor(X, dtype=torch.float64)
    P = torch.linalg.pinv(Xd)                                   # [K, B]
    G = torch.as_tensor(np.asarray(got_w), dtype=torch.float64) @ P      # [H, B] per-sample dh1 (kernel)
    R = torch.as_tensor(np.asarray(ref_w), dtype=torch.float64) @ P      # [H, B] per-sample dh1 (model)
    per = (G - R).norm(dim=0) / R.norm(dim=0).clamp_min(1e-30)
    bad = per > thresh
    keep = ~bad
    clean_w = float(((G - R)[:, keep] @ Xd[keep]).norm() / (R[:, keep] @ Xd[keep]).norm())
    db = torch.as_tensor(np.asarray(got_b), dtype=torch.float64) - torch.as_tensor(np.asarray(ref_b), dtype=torch.float64)
    db_clean = db - (G - R)[:, bad].sum(dim=1)
    clean_b = float(db_clean.norm() / R[:, keep].sum(dim=1).norm())
    return int(bad.sum()), clean_w, clean_b


def rows_off(got, ref, thresh=1e-3):
    """(number of rows whose norm-wise error exceeds ``thresh``, norm-wise error over the other rows): a flipped ReLU unit
    (sample, n) of layer l changes row n of dW_l."""
    got = torch.as_tensor(np.asarray(got), dtype=torch.float64)
    ref = torch.as_tensor(np.asarray(ref), dtype=torch.float64)
    if got.dim() == 1:
        # a bias gradient: one element per unit; measure each against the vector's RMS (an element near zero has no
        # meaningful relative error of its own)
        rows = (got - ref).abs() / ref.pow(2).mean().sqrt().clamp_min(1e-30)
        got, ref = got[:, None], ref[:, None]
    else:
        rows = (got - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-30)
    ok = rows <= thresh
    rest = float((got - ref)[ok].norm() / ref[ok].norm()) if ok.any() else 0.0
    return int((~ok).sum()), rest
