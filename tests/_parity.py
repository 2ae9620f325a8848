"""Helpers shared by the parity tests."""
import torch


def check_full_ranking(got, ref, tol):
    """Full top-k ordering check, k = all classes (BASELINE.json north_star: "identical top-k predicate
    rankings"): walk the REFERENCE ranking of every row; each adjacent pair of that ranking whose gap exceeds
    `tol` must keep its order in `got`.  Returns (decided pairs, all adjacent pairs)."""
    got, ref = got.float().cpu(), ref.float().cpu()
    order = ref.argsort(dim=1, descending=True, stable=True)
    r = ref.gather(1, order)
    g = got.gather(1, order)
    decided = (r[:, :-1] - r[:, 1:]) > tol
    kept = g[:, :-1] > g[:, 1:]
    bad = decided & ~kept
    if bad.any():
        row, col = bad.nonzero()[0].tolist()
        raise AssertionError("ranking differs at row %d, ranks %d/%d: ref %s got %s" % (
            row, col, col + 1, r[row, col:col + 2].tolist(), g[row, col:col + 2].tolist()))
    return int(decided.sum()), decided.numel()


def rel_l2(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return (got - ref).norm().item() / max(ref.norm().item(), 1e-12)
