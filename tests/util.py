"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def legal_move(occ, x, y, dx, dy):
    """scripts/jps1.py:14-31 `not blocked(x, y, dx, dy)` on a uint8 occupancy array (1 = obstacle)."""
    W, H = occ.shape
    tx, ty = x + dx, y + dy
    if tx < 0 or tx >= W or ty < 0 or ty >= H:
        return False
    if dx != 0 and dy != 0:
        if occ[tx, y] == 1 and occ[x, ty] == 1:
            return False
        return occ[tx, ty] != 1
    return occ[tx, ty] != 1


def validate_path(occ, path, start, goal):
    """Every consecutive pair is a straight 8-direction run of legal unit moves.  Returns (straight, diagonal) step counts."""
    assert tuple(path[0]) == tuple(start) and tuple(path[-1]) == tuple(goal), (path[0], path[-1], start, goal)
    a = b = 0
    for (x0, y0), (x1, y1) in zip(path[:-1], path[1:]):
        dx, dy = x1 - x0, y1 - y0
        assert (dx, dy) != (0, 0)
        assert dx == 0 or dy == 0 or abs(dx) == abs(dy), ("not an 8-direction run", (x0, y0), (x1, y1))
        sx, sy = int(np.sign(dx)), int(np.sign(dy))
        n = max(abs(dx), abs(dy))
        x, y = x0, y0
        for _ in range(n):
            assert legal_move(occ, x, y, sx, sy), ("illegal move", (x, y), (sx, sy))
            x, y = x + sx, y + sy
        if sx != 0 and sy != 0:
            b += n
        else:
            a += n
    return a, b


def random_queries(m, n, rng):
    free = np.argwhere(m == 0)
    s = free[rng.integers(len(free), size=n)].astype(np.int32)
    g = free[rng.integers(len(free), size=n)].astype(np.int32)
    return s, g
