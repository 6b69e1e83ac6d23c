"""CPU property checks of the two pieces of mathematics the tensor-core FLAT path (tensor_path.cu) rests on.
No GPU: numpy restatements of what the kernels compute.

1. Threshold sharing.  Every (CTA, query) list publishes the j-th smallest of 32 group minima taken over ANY subset
   of its entries (the kernel uses the most recent <= 64 or <= 256); with j = ceil(K'/slabs), the maximum of those
   values over the slabs is an upper bound of the global K'-th smallest score, so gating on it never drops a row
   of the true top-K'.
2. The margin proof.  With operands rounded to bf16 and products accumulated in fp32, |approx - exact| <= e with
   e = err_coef * |q| * max|x| (+ the kernel's absolute slack); hence if approx[K'] > approx[k] + 2e no row outside
   the K' survivors can belong to the exact top-k."""
import numpy as np
import pytest


def _bf16_round(a):
    """fp32 -> bf16 -> fp32 with round-to-nearest-even (what __float2bfloat16_rn does), vectorised."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("seed", range(6))
def test_published_bound_dominates_global_kth(seed):
    rng = np.random.default_rng(seed)
    slabs = int(rng.integers(1, 149))
    kprime = int(rng.choice([128, 384, 512]))
    j = -(-kprime // slabs)
    if j > 32:
        pytest.skip("the kernel does not publish when j > 32 groups")
    lists = [rng.standard_normal(int(rng.integers(j, 900))).astype(np.float32) for _ in range(slabs)]
    allv = np.sort(np.concatenate(lists))
    if allv.size < kprime:
        pytest.skip("fewer candidates than K'")
    true_kth = allv[kprime - 1]
    bounds = []
    for v in lists:
        win = int(rng.choice([64, 256]))
        tail = v[-win:]                                 # most recent entries
        groups = [tail[g::32] for g in range(32)]       # any partition into 32 groups
        mins = np.sort(np.array([g.min() if g.size else np.inf for g in groups], np.float32))
        bounds.append(mins[j - 1])                      # j-th smallest group minimum (inf if the list is too short)
    assert max(bounds) >= true_kth


@pytest.mark.parametrize("metric,D,scale", [("L2", 768, 1.0), ("IP", 128, 3.0), ("L2", 100, 10.0), ("IP", 1536, 0.05)])
def test_bf16_score_error_is_within_the_kernels_bound(metric, D, scale):
    rng = np.random.default_rng(D)
    N, B = 4000, 8
    X = (scale * rng.standard_normal((N, D))).astype(np.float32)
    Q = (scale * rng.standard_normal((B, D))).astype(np.float32)
    Xh, Qh = _bf16_round(X), _bf16_round(Q)
    # fp32 accumulation of exactly representable bf16 products (the tensor core's accumulation order differs; its
    # extra error is what the 2e-4 term of the coefficient is for) vs the exact score in float64
    dot_h = (Xh.astype(np.float32) @ Qh.astype(np.float32).T).astype(np.float32)
    dot = X.astype(np.float64) @ Q.astype(np.float64).T
    xn32 = np.sum(X.astype(np.float32) ** 2, axis=1, dtype=np.float32)
    xn = np.sum(X.astype(np.float64) ** 2, axis=1)
    if metric == "L2":
        approx = xn32[:, None] - 2.0 * dot_h           # the kernel's score: |x|^2 - 2 x.q  (|q|^2 is per-query constant)
        exact = xn[:, None] - 2.0 * dot
        coef = 2.0 * (0.00390625 * 1.02 + 2e-4)        # RerankParams::err_coef for L2
    else:
        approx = -dot_h
        exact = -dot
        coef = 0.00390625 * 1.02 + 2e-4
    xmax = np.sqrt(xn.max())
    qn = np.sqrt(np.sum(Q.astype(np.float64) ** 2, axis=1))
    e = coef * qn[None, :] * xmax + 1e-5 * xmax * xmax + 1e-30
    assert np.all(np.abs(approx - exact) <= e), float(np.max(np.abs(approx - exact) / e))


def test_margin_rule_keeps_the_exact_topk():
    """If approx[K'] > approx[k] + 2e (approx ascending), every exact top-k row is among the K' best by approx."""
    rng = np.random.default_rng(3)
    N, k, kprime = 20000, 100, 384
    exact = rng.standard_normal(N)
    for e in (0.0005, 0.002, 0.01):
        approx = exact + rng.uniform(-e, e, N)          # any perturbation bounded by e
        order = np.argsort(approx)
        a_sorted = approx[order]
        if a_sorted[kprime - 1] > a_sorted[k - 1] + 2 * e:   # the kernel's acceptance test
            survivors = set(order[:kprime].tolist())
            assert set(np.argsort(exact)[:k].tolist()) <= survivors
