"""Offline search of the shared-memory bank swizzle masks (csrc/dmx_plan.hpp: SWZ_M).

Tile positions are index ^ hash(index >> 4 ...) on the low 4 bits (8-byte bank pairs of a 128-byte line).  A sweep keeps
the codes of 2 or 3 tile qubits in registers; the 16 lanes of a half-warp walk the four lowest remaining index bits.  The
access is conflict-free iff the images of those four bits in the 4-bit bank-pair index are linearly independent over
GF(2).  This script hill-climbs the 4 x 10 hash bits until every (tile size, resident positions) case has rank 4."""
import itertools
import random


def rank(vs):
    vs, r = list(vs), 0
    for b in range(4):
        piv = next((i for i in range(r, len(vs)) if (vs[i] >> b) & 1), None)
        if piv is None:
            continue
        vs[r], vs[piv] = vs[piv], vs[r]
        for i in range(len(vs)):
            if i != r and (vs[i] >> b) & 1:
                vs[i] ^= vs[r]
        r += 1
    return r


def constraints():
    cons = {(1, 2, 3, 4)}                       # linear staging loop: lanes walk bits 1..4 (16-byte chunks)
    for k in (5, 6, 7):
        for nq in (2, 3):
            for pos in itertools.combinations(range(0, 2 * k, 2), nq):
                excl = {b for p in pos for b in (p, p + 1)}
                free = [b for b in range(2 * k) if b not in excl]
                if len(free) >= 4:
                    cons.add(tuple(free[:4]))
    return sorted(cons)


def main(seed=1):
    cons = constraints()
    cost = lambda cols: sum(rank([(1 << b) ^ cols[b] if b < 4 else cols[b] for b in f]) < 4 for f in cons)
    random.seed(seed)
    while True:
        cols = [0] * 4 + [random.randrange(16) for _ in range(10)]
        c = cost(cols)
        for _ in range(400):
            if c == 0:
                break
            b = random.randrange(4, 14)
            old, cols[b] = cols[b], random.randrange(16)
            c2 = cost(cols)
            if c2 <= c:
                c = c2
            else:
                cols[b] = old
        if c == 0:
            masks = [sum(1 << b for b in range(4, 14) if (cols[b] >> i) & 1) for i in range(4)]
            print("SWZ_M =", [hex(m) for m in masks])
            return masks


if __name__ == "__main__":
    main()
