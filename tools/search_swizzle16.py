"""Feasibility study for DESIGN.md §9 (NOT used by the engine): a bank swizzle for 16-byte tile exchanges.

With a pivot bit p (a register bit shared by the storing and the loading sweep) moved to address bit 0, a thread moves
register pairs as 16-byte units; unit index = the tile index with bit p removed.  A quarter-warp (8 lanes) of LDS.128 /
STS.128 is one 128-byte wavefront, conflict-free iff the 8 lanes hit the 8 distinct 16-byte bank groups, i.e. iff the
images of the three lowest lane-walked unit-index bits in the 3-bit group index are linearly independent over GF(2).
Position = unit ^ hash(unit >> 3) on the low 3 bits; this script hill-climbs the 3 x (2k - 4) hash bits looking for one
STATIC hash under which every (tile size k, resident triple, pivot bit) case has rank 3.

Result (round 1): none exists.  The 18 distinct lane patterns include (0,1,j) and (1,2,j) for j = 3, 5, 7, which force the
images of unit bits 3, 5 and 7 into {101, 111}, while (3,4,5), (3,4,7), (4,5,7), (5,6,7) need them pairwise distinct — three
bits, two values.  The best hash leaves 3 of 18 patterns with 2-way conflicts.  So a 16-byte exchange needs the swizzle (or
the order of the non-pivot bits in the unit index) chosen PER EXCHANGE — possible, because the tile addressing is
table-driven per sweep (masks precombined on the host), at the price of per-sweep hash masks for the thread base."""
import itertools
import random


def rank3(vs):
    vs, r = list(vs), 0
    for b in range(3):
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
    cons = set()
    for k in (6, 7):
        for pos in itertools.combinations(range(0, 2 * k, 2), 3):
            resident = sorted(b for p in pos for b in (p, p + 1))
            for pivot in resident:
                unit_of = {}                                    # tile bit -> unit-index bit (pivot removed)
                for b in range(2 * k):
                    if b != pivot:
                        unit_of[b] = len(unit_of)
                free = [unit_of[b] for b in range(2 * k) if b not in resident]
                cons.add(tuple(free[:3]))
    return sorted(cons)


def main(seed=1, n_hash_bits=10):
    cons = constraints()
    image = lambda cols, b: ((1 << b) ^ cols[b]) if b < 3 else cols[b]
    cost = lambda cols: sum(rank3([image(cols, b) for b in f]) < 3 for f in cons)
    random.seed(seed)
    for attempt in range(200):
        cols = [0] * 3 + [random.randrange(8) for _ in range(n_hash_bits)]
        c = cost(cols)
        for _ in range(600):
            if c == 0:
                break
            b = random.randrange(3, 3 + n_hash_bits)
            old, cols[b] = cols[b], random.randrange(8)
            c2 = cost(cols)
            if c2 <= c:
                c = c2
            else:
                cols[b] = old
        if c == 0:
            masks = [sum(1 << b for b in range(3, 3 + n_hash_bits) if (cols[b] >> i) & 1) for i in range(3)]
            print(f"{len(cons)} lane patterns conflict-free; group-index hash masks (unit-index bits):", [hex(m) for m in masks])
            return masks
    print("no static swizzle serves all", len(cons), "lane patterns (see the module docstring)")
    return None


if __name__ == "__main__":
    main()
