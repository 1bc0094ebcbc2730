"""The claim behind the table of bracket points in the REML fit kernel (cellregmap_b200/csrc/fit.cuh: FIT_TAB_*, fit_bracket_point):
whatever the objective, the bracket search that LMM.fit runs on logit(delta) visits a prefix of one of two fixed point sequences.
Checked on the oracle's restatement of brent-search (oracle/brent_port.py) -- no device involved."""
import math

import numpy as np

from oracle import brent_port

LOGMAX = 709.782712893384


def bracket_point(direction, k):
    """Python twin of fit_bracket_point (fit.cuh): point k of the sequence towards +inf (direction 0) or -inf (direction 1)."""
    a0, b0, rtol, atol, gfactor = -LOGMAX, LOGMAX, 1e-6, 1e-6, 2.0
    x0 = min(max(0.0, a0), b0)
    step0 = gfactor * (rtol * abs(x0) + atol)
    x1 = max(x0 - step0, a0) if x0 - a0 > b0 - x0 else min(x0 + step0, b0)
    if k == 0:
        return x0
    if k == 1:
        return x1
    if direction:
        x0, x1 = x1, x0
    for _ in range(2, k + 1):
        x2 = min(max(x1 + (x1 - x0) * gfactor, a0), b0)
        x0, x1 = x1, x2
    return x1


def test_bracket_walks_one_of_two_fixed_sequences():
    rng = np.random.default_rng(0)
    seen_dirs = set()
    longest = 0
    for trial in range(400):
        centre = rng.uniform(-30, 30) if trial % 4 else rng.choice([-LOGMAX * 2, LOGMAX * 2, 0.0, 1e-7])
        width = 10.0 ** rng.uniform(-3, 2)
        kind = trial % 3

        def f(x, centre=centre, width=width, kind=kind):
            z = (x - centre) / width
            if kind == 0:
                return z * z
            if kind == 1:
                return math.log1p(z * z) + 0.01 * math.sin(3.0 * x)
            return abs(z) ** 1.5 - 0.3 * math.exp(-z * z)

        visited = []

        def recording(x):
            visited.append(x)
            return f(x)

        brent_port.bracket(recording, a=-LOGMAX, b=LOGMAX, rtol=1e-6, atol=1e-6)
        assert visited[0] == bracket_point(0, 0) and visited[1] == bracket_point(0, 1)
        if len(visited) > 2:
            direction = 0 if visited[2] > visited[1] else 1
            seen_dirs.add(direction)
            for k, x in enumerate(visited):
                if k >= 2:
                    assert x == bracket_point(direction, k), (trial, k, x)      # bit for bit
        longest = max(longest, len(visited))
    assert seen_dirs == {0, 1}
    assert longest <= 30 + 2            # FIT_TAB_K = 30 points per direction cover every search (2e-6 * 2^30 >> LOGMAX is hit first)


def test_sequences_end_at_the_bounds():
    for direction in (0, 1):
        xs = [bracket_point(direction, k) for k in range(34)]
        bound = LOGMAX if direction == 0 else -LOGMAX
        assert xs[-1] == bound and xs[-2] == bound
        first_at_bound = xs.index(bound)
        assert first_at_bound <= 30
        inner = xs[2:first_at_bound]
        assert all(abs(b) > abs(a) for a, b in zip(inner, inner[1:]))
