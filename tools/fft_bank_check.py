"""Shared-memory bank check of the radix-8 Stockham passes in csrc/wb_fft.h.

Every access is a 16-byte wb_cplx (LDS.128 / STS.128): the hardware serves a warp in four phases of 8 lanes, and a
phase is conflict-free when its 8 addresses fall into 8 distinct 16-byte bank groups (slot mod 8).  For every transform
size the fast path handles, every pass and every register index q / m, this script replays the slot each lane
touches (with the XOR swizzle slot = i ^ ((i >> 3) & 7) on the intermediate buffers) and prints the worst conflict
degree.  Expected output: degree 1 everywhere.
"""


def swz(i):
    return i ^ ((i >> 3) & 7)


def plan(n):
    ln = n.bit_length() - 1
    r0 = 1 << (ln % 3)
    return r0, ln // 3


def degree(slots):
    worst = 1
    for p in range(0, len(slots), 8):
        grp = slots[p:p + 8]
        banks = {}
        for s in grp:
            banks.setdefault(s % 8, set()).add(s)
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst


def check(n, nthr):
    r0, p = plan(n)
    passes = []
    ns = 1
    if r0 > 1:
        passes.append((r0, 1))
        ns = r0
    for _ in range(p):
        passes.append((8, ns))
        ns *= 8
    out = []
    for idx, (r, ns) in enumerate(passes):
        tb = n // r
        swz_in = idx > 0
        swz_out = idx + 1 < len(passes)
        worst_r = worst_w = 1
        for j0 in range(0, tb, 32):
            lanes = [j for j in range(j0, min(tb, j0 + 32))]
            for q in range(r):
                rd = [(swz(j + q * tb) if swz_in else j + q * tb) for j in lanes]
                worst_r = max(worst_r, degree(rd))
                wr = []
                for j in lanes:
                    k = j & (ns - 1)
                    o = (j - k) * r + k + q * ns
                    wr.append(swz(o) if swz_out else o)
                worst_w = max(worst_w, degree(wr))
        out.append((r, ns, worst_r, worst_w))
    return out


if __name__ == "__main__":
    bad = 0
    for n in (64, 128, 256, 512, 1024, 2048, 4096):
        for r, ns, wr, ww in check(n, 256):
            print("N=%5d radix-%d stride %4d: read conflict degree %d, write conflict degree %d" % (n, r, ns, wr, ww))
            bad += (wr > 1) + (ww > 1)
    print("OK" if not bad else "%d conflicting access patterns" % bad)
