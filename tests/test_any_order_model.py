"""The exactness argument of the any-order walk as an executable model — no GPU, no oracle.

A ray sees leaves p = 0..n-1 in the reference's depth-first order, each with a computed hit distance t_p and
m_p = max(t_p, entry of its own box) >= t_p (m_p > t_p: an ABNORMAL leaf). The reference is the sequential process
    T = t_max; for p in order: if T >= m_p: T = t_p; best = p
The any-order walk looks at the leaves in an arbitrary order, skips a leaf whose box entry lies above the current window
top T_win = min(t_max, t_best + 2 slack), tests the others with T_win as the upper bound, keeps min t (ties: larger rank) over
NORMAL leaves, remembers the smallest abnormal t it tested, and falls back to the sequential process when that t is inside
the final window. With m_p - t_p <= slack for every leaf that can be skipped, both must agree on every input."""
import random

import pytest


def sequential(leaves, t_max):
    T, best = t_max, None
    for p, (t, m, _near) in enumerate(leaves):
        if T >= m:
            T, best = t, p
    return best, (T if best is not None else None)


def any_order(leaves, t_max, slack, order, big):
    """returns (best, t, fell_back)"""
    T, best = t_max, None
    T_win = t_max
    a_min = None
    for p in order:
        t, m, near = leaves[p]
        if p not in big and near > T_win:
            continue  # box beyond the window: skipped without a test (a big leaf is never skipped)
        if t > T_win:
            continue  # primitive test rejects
        if m == t:  # normal
            if best is None or t < T or (t == T and p > best):
                T, best = t, p
                T_win = min(T_win, t + 2 * slack)
        else:
            a_min = t if a_min is None else min(a_min, t)
    if a_min is not None and a_min <= T_win:
        b, t = sequential(leaves, t_max)
        return b, t, True
    return best, (T if best is not None else None), False


def random_case(rng, n, slack, p_abnormal, grid):
    leaves, big = [], set()
    for p in range(n):
        t = rng.randrange(0, grid)
        if rng.random() < 0.1:
            big.add(p)
            gap = rng.randrange(0, 4 * slack + 3) if rng.random() < p_abnormal else 0  # no bound on a big leaf's gap
        else:
            gap = rng.randrange(1, slack + 1) if (slack > 0 and rng.random() < p_abnormal) else 0
        m = t + gap
        near = m if gap else rng.randrange(0, t + 1)  # a normal leaf's box is entered at or before its hit
        leaves.append((t, m, near))
    t_max = rng.choice([grid + 10, grid + 10, rng.randrange(0, grid)])
    return leaves, big, t_max


@pytest.mark.parametrize("slack,p_abnormal", [(0, 0.0), (1, 0.05), (1, 0.5), (3, 0.2), (5, 0.9)])
def test_any_order_model_equals_sequential_process(slack, p_abnormal):
    rng = random.Random(1234 + slack)
    fallbacks = 0
    for case in range(6000):
        n = rng.randrange(1, 14)
        leaves, big, t_max = random_case(rng, n, slack, p_abnormal, grid=12)  # a coarse grid forces exact ties
        want = sequential(leaves, t_max)
        for _ in range(3):
            order = list(range(n))
            rng.shuffle(order)
            b, t, fb = any_order(leaves, t_max, slack, order, big)
            fallbacks += fb
            assert (b, t) == want, (leaves, big, t_max, order, (b, t), want)
    if p_abnormal == 0.0:
        assert fallbacks == 0


def test_any_order_model_needs_the_window():
    """the same walk with a window of only ONE slack is wrong: a skipped abnormal leaf can block the winner"""
    # leaf 0 is abnormal (t = 5, box entry 6), leaf 1 is normal with t = 5.5 -> the reference accepts 0 (T = inf >= 6) and then
    # rejects 1 (5 < 5.5); a walk that tests leaf 1 first and then skips leaf 0 because 6 > 5.5 + 0 would answer 1
    leaves = [(10, 12, 12), (11, 11, 3)]
    assert sequential(leaves, 100) == (0, 10)
    b, t, fb = any_order(leaves, 100, 2, [1, 0], set())
    assert (b, t) == (0, 10) and fb  # window 11 + 4 keeps leaf 0 in sight, it is abnormal inside the window: sequential fallback
    b, t, fb = any_order(leaves, 100, 0, [1, 0], set())  # slack understated: leaf 0 is skipped and the answer is wrong
    assert (b, t) == (1, 11) and not fb
