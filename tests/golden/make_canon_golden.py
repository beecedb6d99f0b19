#!/usr/bin/env python
"""Golden table of the reference's successor rule, generated from the UNMODIFIED scripts/jps1.py.

Run in the build container only (needs /root/reference):   python tests/golden/make_canon_golden.py
Writes canon_succ_golden.json next to this file: for every occupancy pattern of the 8 cells around a free centre
cell (256 patterns, in a 3x3 array so that the array border never matters, and again in a 1-cell-wide corner
position so that `blocked` == True outside the array is exercised) and every arrival direction (8 parents + the
start), the set of neighbours `jps1.nodeNeighbours` returns (jps1.py:49-93), filtered by `not jps1.blocked` for the
move (jps1.jump rejects the others at :99 / by the legality of the step), as a bit mask over the direction order of
include/fuxi_b200.h, together with the legal-move mask of the centre cell.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

DIRS = [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (-1, 1), (1, -1), (1, 1)]


def case(j, m, cx, cy):
    W, H = m.shape
    moves = 0
    for d, (dx, dy) in enumerate(DIRS):
        if not j.blocked(cx, cy, dx, dy, m):
            moves |= 1 << d
    out = []
    for code in range(9):
        if code < 8:
            px, py = cx - DIRS[code][0], cy - DIRS[code][1]
            if not (0 <= px < W and 0 <= py < H):
                out.append(None)      # no such parent inside the array
                continue
            # the parent must have been able to make the move (otherwise the cell is never reached this way)
            if j.blocked(px, py, DIRS[code][0], DIRS[code][1], m):
                out.append(None)
                continue
            nb = j.nodeNeighbours(cx, cy, (px, py), m)
        else:
            nb = j.nodeNeighbours(cx, cy, 0, m)
        succ = 0
        for (nx, ny) in nb:
            dx, dy = nx - cx, ny - cy
            if not j.blocked(cx, cy, dx, dy, m):
                succ |= 1 << DIRS.index((dx, dy))
        out.append(succ)
    return moves, out


def main():
    assert refload.available(), "reference tree not found"
    j = refload.jps1_module()
    recs = []
    # (a) centre of a 5x5 array: all 8 neighbours and all parents exist; the outer ring is free
    for pat in range(256):
        m = np.zeros((5, 5))
        for d, (dx, dy) in enumerate(DIRS):
            if (pat >> d) & 1:
                m[2 + dx][2 + dy] = 1
        moves, out = case(j, m, 2, 2)
        recs.append({"where": "interior", "pattern": pat, "moves": moves, "succ": out})
    # (b) cells on the border / in the corner of a 3x3 array: outside counts as blocked
    for (cx, cy) in [(0, 0), (0, 1), (1, 0), (2, 2), (2, 1), (1, 2), (0, 2), (2, 0)]:
        for pat in range(256):
            m = np.zeros((3, 3))
            ok = True
            for d, (dx, dy) in enumerate(DIRS):
                x, y = cx + dx, cy + dy
                if (pat >> d) & 1:
                    if 0 <= x < 3 and 0 <= y < 3:
                        m[x][y] = 1
                    else:
                        ok = False      # pattern bit on a cell that does not exist: skip (covered by another pattern)
            if not ok:
                continue
            moves, out = case(j, m, cx, cy)
            recs.append({"where": "border%d%d" % (cx, cy), "pattern": pat, "moves": moves, "succ": out})
    with open(os.path.join(HERE, "canon_succ_golden.json"), "w") as f:
        json.dump({"dirs": DIRS, "records": recs}, f, separators=(",", ":"))
    print("wrote", len(recs), "records")


if __name__ == "__main__":
    main()
