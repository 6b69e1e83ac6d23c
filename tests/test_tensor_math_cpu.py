"""CPU property checks of the two pieces of mathematics the tensor-core FLAT path (tensor_path.cu) rests on.
No GPU: numpy restatements of what the kernels compute.

1. Threshold sharing.  Every (CTA, query) list publishes the j-th smallest of 32 group minima taken over ANY subset
   of its entries (the kernel uses the most recent <= 64 or <= 256); with j = ceil(K'/slabs), the maximum of those
   values over the slabs is an upper bound of the global K'-th smallest score, so gating on it never drops a row
   of the true top-K'.
2. The margin proof.  With operands rounded to bf16 and products accumulated in fp32, |approx - exact| <= e with
   e = err_coef * |q| * max|x| (+ the kernel's absolute slack); the survivors are re-ranked exactly, and if
   approx[K'] - e > (exact k-th best among the survivors) no row outside the K' survivors can belong to the exact
   top-k (rerank_kernel; the fp32 rounding of the reference's own distance is a further (1 - rho) on the left)."""
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
    """Acceptance test of rerank_kernel: gK - e > dk, with gK the K'-th smallest approximate score and dk the exact
    k-th best among the K' survivors.  Whenever it holds, the exact top-k of ALL rows is the exact top-k of the
    survivors — for random and for adversarial perturbations bounded by e."""
    rng = np.random.default_rng(3)
    N, k, kprime = 20000, 100, 384
    accepted = 0
    for trial in range(60):
        exact = rng.standard_normal(N)
        e = float(rng.choice([0.0005, 0.002, 0.01, 0.03]))
        if trial % 3 == 0:
            approx = exact + rng.uniform(-e, e, N)
        elif trial % 3 == 1:   # adversarial: the best rows look as bad as allowed, the next ones as good as allowed
            approx = exact.copy()
            order_x = np.argsort(exact)
            approx[order_x[:k]] += e
            approx[order_x[k:]] -= e
        else:                  # extreme values only
            approx = exact + e * rng.choice([-1.0, 1.0], N)
        order = np.argsort(approx)
        surv = order[:kprime]
        gK = approx[order[kprime - 1]]
        dk = np.sort(exact[surv])[k - 1]
        if gK - e > dk:
            accepted += 1
            want = np.argsort(exact)[:k]
            got = surv[np.argsort(exact[surv])[:k]]
            assert np.array_equal(np.sort(want), np.sort(got))
            # strictly: every outsider is beyond dk
            outsiders = order[kprime:]
            assert exact[outsiders].min() > dk
    assert accepted >= 10


def test_new_rule_accepts_whatever_the_old_rule_accepted():
    """gK > gk + 2e (the round-1 test on approximate scores alone) implies gK - e > dk: the re-ranked k-th distance is
    at most gk + e.  The new rule flags fewer queries for the exact-scan fallback, never more."""
    rng = np.random.default_rng(9)
    N, k, kprime = 5000, 50, 214
    for _ in range(200):
        exact = rng.standard_normal(N)
        e = float(rng.choice([0.001, 0.01, 0.05]))
        approx = exact + rng.uniform(-e, e, N)
        order = np.argsort(approx)
        gk, gK = approx[order[k - 1]], approx[order[kprime - 1]]
        dk = np.sort(exact[order[:kprime]])[k - 1]
        if gK > gk + 2 * e:
            assert gK - e > dk
