"""Flop model of block elimination orderings for the batched block LDL^T (Ned_RT, n fine cells per direction).

Restates the symbolic phase of csrc/topology.cpp:build_direct_plan (block graph, fill, 32-padding, panel-wise
trailing-update flops incl. the 32 rhs rows) so that candidate orderings can be compared without a GPU.
Usage: python profiles/ordering_model.py [n]
"""
import itertools
import sys
from collections import defaultdict

import numpy as np

PW = 32
RHS = 32


def dofs_ned_rt(n):
    """interior DoFs: (pos*2 as ints, is_u).  edges first (sigma), faces (u)."""
    out = []
    for ax in range(3):
        for a in range(n):              # along the edge
            for b in range(1, n):
                for c in range(1, n):
                    p = [0, 0, 0]
                    p[ax] = 2 * a + 1
                    p[(ax + 1) % 3] = 2 * b
                    p[(ax + 2) % 3] = 2 * c
                    out.append((tuple(p), 0))
    for ax in range(3):
        for a in range(1, n):           # normal coordinate
            for b in range(n):
                for c in range(n):
                    p = [0, 0, 0]
                    p[ax] = 2 * a
                    p[(ax + 1) % 3] = 2 * b + 1
                    p[(ax + 2) % 3] = 2 * c + 1
                    out.append((tuple(p), 1))
    return out


def cells_of(p, n):
    rng = []
    for d in range(3):
        if p[d] % 2 == 1:
            rng.append([(p[d] - 1) // 2])
        else:
            rng.append([c for c in (p[d] // 2 - 1, p[d] // 2) if 0 <= c < n])
    return list(itertools.product(*rng))


def cost(blocks, dofs, n, verbose=False):
    """blocks: list of lists of dof indices in elimination order."""
    nB = len(blocks)
    blk_of = {}
    for s, b in enumerate(blocks):
        for d in b:
            blk_of[d] = s
    assert len(blk_of) == len(dofs)
    cell_blocks = defaultdict(set)
    for i, (p, _) in enumerate(dofs):
        for c in cells_of(p, n):
            cell_blocks[c].add(blk_of[i])
    adj = np.zeros((nB, nB), bool)
    for bs in cell_blocks.values():
        for a in bs:
            for b in bs:
                if a < b:
                    adj[a, b] = True
    size = [(len(b) + PW - 1) // PW * PW for b in blocks]
    flops = 0.0
    band = 0
    panels = 0
    maxfront = 0
    for s in range(nB):
        reach = [b for b in range(s + 1, nB) if adj[s, b]]
        for i in range(len(reach)):
            for j in range(i + 1, len(reach)):
                adj[reach[i], reach[j]] = True
        rows = size[s] + sum(size[b] for b in reach)
        ld = rows + RHS
        band += ld * size[s]
        maxfront = max(maxfront, rows)
        for j0 in range(0, size[s], PW):
            R = ld - (j0 + PW)
            Cn = rows - (j0 + PW)
            panels += 1
            if Cn > 0:
                flops += 2.0 * PW * (Cn * R - Cn * (Cn - 1) / 2.0)
        if verbose:
            print(f"  block {s}: {len(blocks[s])} -> {size[s]}, front {rows}, reach {reach}")
    return dict(gflop=flops / 1e9, band_mb=band * 8 / 1e6, blocks=nB, panels=panels, maxfront=maxfront,
                NP=sum(size))


def order_sigma_first(idx, dofs):
    return sorted(idx, key=lambda i: (dofs[i][1], dofs[i][0][2], dofs[i][0][1], dofs[i][0][0]))


def layers_planes(dofs, n):
    by = defaultdict(list)
    for i, (p, _) in enumerate(dofs):
        z = p[2]
        key = z - 1 if z % 2 == 0 else z - 1  # plane z=2k -> 2k-1 ; layer z=2k+1 -> 2k
        by[key].append(i)
    return [order_sigma_first(by[k], dofs) for k in sorted(by)]


def nd(dofs, n, box, axes_order, min_cells, sep_split):
    """recursive geometric nested dissection of box = [(lo,hi)]*3 in doubled coordinates (open interior).
    returns list of blocks.  sep_split: split separators by the child separators' planes (finer blocks)."""
    idx_all = [i for i, (p, _) in enumerate(dofs) if all(box[d][0] < p[d] < box[d][1] for d in range(3))]

    def rec(box, idx, depth):
        ext = [(box[d][1] - box[d][0]) // 2 for d in range(3)]
        cand = [d for d in axes_order if ext[d] > min_cells]
        if not cand or len(idx) <= PW:
            return [idx] if idx else []
        d = max(cand, key=lambda a: (ext[a], -axes_order.index(a)))
        mid = box[d][0] + (ext[d] // 2) * 2
        left = [i for i in idx if dofs[i][0][d] < mid]
        right = [i for i in idx if dofs[i][0][d] > mid]
        sep = [i for i in idx if dofs[i][0][d] == mid]
        bl = list(box); bl[d] = (box[d][0], mid)
        br = list(box); br[d] = (mid, box[d][1])
        out = rec(bl, left, depth + 1) + rec(br, right, depth + 1)
        if sep_split:
            bs = list(box); bs[d] = (mid - 1, mid + 1)
            # dissect the separator plane itself with the remaining axes
            out += rec_sep(bs, sep, d)
        else:
            out.append(sep)
        return out

    def rec_sep(box, idx, fixed):
        ext = [(box[d][1] - box[d][0]) // 2 for d in range(3)]
        cand = [d for d in range(3) if d != fixed and ext[d] > min_cells]
        if not cand or len(idx) <= 2 * PW:
            return [idx] if idx else []
        d = max(cand, key=lambda a: ext[a])
        mid = box[d][0] + (ext[d] // 2) * 2
        left = [i for i in idx if dofs[i][0][d] < mid]
        right = [i for i in idx if dofs[i][0][d] > mid]
        sep = [i for i in idx if dofs[i][0][d] == mid]
        bl = list(box); bl[d] = (box[d][0], mid)
        br = list(box); br[d] = (mid, box[d][1])
        return rec_sep(bl, left, fixed) + rec_sep(br, right, fixed) + ([sep] if sep else [])

    blocks = rec(box, idx_all, 0)
    return [order_sigma_first(b, dofs) for b in blocks if b]


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dofs = dofs_ned_rt(n)
    print("interior dofs", len(dofs))
    print("layers/planes", cost(layers_planes(dofs, n), dofs, n))
    box = [(0, 2 * n)] * 3
    for min_cells in (4, 2, 1):
        for sep_split in (False, True):
            for axes in ([2, 1, 0],):
                b = nd(dofs, n, box, axes, min_cells, sep_split)
                print(f"nd min_cells={min_cells} sep_split={sep_split}", cost(b, dofs, n))


def deferral(dofs, n, keep):
    """layers/planes with u-type DoFs deferred to the next block so that block sizes are multiples of 32.
    keep(size) -> number of DoFs the block keeps."""
    blocks = layers_planes(dofs, n)
    out = []
    carry = []
    for b in blocks:
        cur = carry + b
        cur = order_sigma_first(cur, dofs)
        k = keep(len(cur))
        if b is blocks[-1]:
            k = len(cur)
        # defer the last u-type DoFs
        n_u = sum(dofs[i][1] for i in cur)
        d = min(len(cur) - k, n_u)
        carry = cur[len(cur) - d:] if d > 0 else []
        out.append(cur[:len(cur) - d] if d > 0 else cur)
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dofs = dofs_ned_rt(n)
    print("current (defer if excess <= 8):", cost(deferral(dofs, n, lambda s: s - s % 32 if 0 < s % 32 <= 8 else s), dofs, n))
    print("defer down to multiple of 32 always:", cost(deferral(dofs, n, lambda s: s - s % 32), dofs, n))
    print("defer if excess <= 16:", cost(deferral(dofs, n, lambda s: s - s % 32 if 0 < s % 32 <= 16 else s), dofs, n))
    print("defer if excess <= 20:", cost(deferral(dofs, n, lambda s: s - s % 32 if 0 < s % 32 <= 20 else s), dofs, n))
