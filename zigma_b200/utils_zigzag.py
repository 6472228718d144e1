"""Scan-path tables of ZigMa (``utils/utils_zigzag.py`` of the reference): integer, bit exact.

``zigzag_path(N)`` (:144-175) -- 8 boustrophedon paths over the N x N token grid: for every start
corner a row snake and a column snake.  ``reverse_permut_np`` (:136-141).  ``hilbert_path(N)``
(:285-302) -- the 8 rot90/transpose variants of the generalised-Hilbert ORDER INDEX map (the
reference gathers with the order-index array itself; kept as is).  Vectorised numpy; no plotting
dependencies (the reference imports matplotlib at module import, :4-5).
"""
import numpy as np


def zigzag_path(N):
    line, pos = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    snake = np.where(line % 2 == 0, pos, N - 1 - pos)      # boustrophedon position inside a line
    paths = []
    for r0, c0, dr, dc in ((0, 0, 1, 1), (0, N - 1, 1, -1), (N - 1, 0, -1, 1), (N - 1, N - 1, -1, -1)):
        paths.append(((r0 + dr * line) * N + (c0 + dc * snake)).reshape(-1).astype(np.int64))   # rows
        paths.append(((r0 + dr * snake) * N + (c0 + dc * line)).reshape(-1).astype(np.int64))   # columns
    return paths


def reverse_permut_np(permutation):
    permutation = np.asarray(permutation)
    reverse = np.empty(len(permutation), dtype=np.int64)
    reverse[permutation] = np.arange(len(permutation), dtype=np.int64)
    return reverse


def _gilbert_index(x, y, w, h):
    """Position of cell (x, y) along the generalised Hilbert curve of a w x h grid
    (utils_zigzag.py:16-120), written as a loop over the subdivision levels."""
    sgn = lambda v: (v > 0) - (v < 0)

    def contains(qx, qy, ox, oy, ax, ay, bx, by):
        dx, dy = ax + bx, ay + by
        in_x = (ox + dx < qx <= ox) if dx < 0 else (ox <= qx < ox + dx)
        in_y = (oy + dy < qy <= oy) if dy < 0 else (oy <= qy < oy + dy)
        return in_x and in_y

    idx, ox, oy = 0, 0, 0
    ax, ay, bx, by = (w, 0, 0, h) if w >= h else (0, h, w, 0)
    while True:
        wid, hei = abs(ax + ay), abs(bx + by)
        dax, day, dbx, dby = sgn(ax), sgn(ay), sgn(bx), sgn(by)
        if hei == 1 or wid == 1:
            along_y = (dax == 0) if hei == 1 else (dbx == 0)
            return idx + ((day + dby) * (y - oy) if along_y else (dax + dbx) * (x - ox))
        ax2, ay2, bx2, by2 = ax // 2, ay // 2, bx // 2, by // 2
        if 2 * wid > 3 * hei:                       # long strip: split the major axis only
            if abs(ax2 + ay2) % 2 and wid > 2:
                ax2, ay2 = ax2 + dax, ay2 + day
            if contains(x, y, ox, oy, ax2, ay2, bx, by):
                ax, ay = ax2, ay2
            else:
                idx += abs((ax2 + ay2) * (bx + by))
                ox, oy, ax, ay = ox + ax2, oy + ay2, ax - ax2, ay - ay2
            continue
        if abs(bx2 + by2) % 2 and hei > 2:
            bx2, by2 = bx2 + dbx, by2 + dby
        if contains(x, y, ox, oy, bx2, by2, ax2, ay2):          # step up
            ax, ay, bx, by = bx2, by2, ax2, ay2
            continue
        idx += abs((bx2 + by2) * (ax2 + ay2))
        if contains(x, y, ox + bx2, oy + by2, ax, ay, bx - bx2, by - by2):   # long horizontal
            ox, oy, bx, by = ox + bx2, oy + by2, bx - bx2, by - by2
            continue
        idx += abs((ax + ay) * ((bx - bx2) + (by - by2)))       # step down
        ox, oy = ox + (ax - dax) + (bx2 - dbx), oy + (ay - day) + (by2 - dby)
        ax, ay, bx, by = -bx2, -by2, -(ax - ax2), -(ay - ay2)


def gilbert_zigzag_path(N):
    order = np.zeros((N, N), dtype=np.int64)
    for x in range(N):
        for y in range(N):
            order[x, y] = _gilbert_index(x, y, N, N)
    return order


def hilbert_path(N=16):
    base = gilbert_zigzag_path(N)
    variants = []
    for k in range(4):
        rot = np.rot90(base, k)
        variants += [rot, rot.T]
    return [np.ascontiguousarray(v).reshape(-1) for v in variants]
