"""Generates universal-volumetric_b200/csrc/uastc_tables.h (partition patterns, anchor masks, endpoint unquantisation).

The patterns come from the ASTC partition hash applied to the seeds of the three UASTC common-partition tables;
the script first re-checks those tables against the BC7 partition tables (the consistency argument that pins
oracle/uastc_oracle.c, see its header): every {BC7 id, ASTC seed, inversion / permutation / merge} triple must hold,
and an exhaustive search over the 1024 seeds must find exactly the same BC7 ids.
Run: python tools/gen/gen_uastc_tables.py   (writes the header in place; deterministic)
"""
import itertools
import os

M = 0xFFFFFFFF


def hash52(p):
    p &= M
    p ^= p >> 15; p = (p - (p << 17)) & M; p = (p + (p << 7)) & M; p = (p + (p << 4)) & M; p ^= p >> 5
    p = (p + (p << 16)) & M; p ^= p >> 7; p ^= p >> 3; p ^= (p << 6) & M; p ^= p >> 17
    return p


def select_partition(seed, x, y, count):
    x <<= 1; y <<= 1                      # 4x4 is a "small block"
    seed += (count - 1) * 1024
    r = hash52(seed)
    s = [(r >> (4 * i)) & 15 for i in range(8)]
    s = [v * v for v in s]
    if seed & 1:
        sh1 = 4 if seed & 2 else 5; sh2 = 6 if count == 3 else 5
    else:
        sh1 = 6 if count == 3 else 5; sh2 = 4 if seed & 2 else 5
    s = [v >> (sh2 if i & 1 else sh1) for i, v in enumerate(s)]
    a = (s[0] * x + s[1] * y + (r >> 14)) & 63
    b = (s[2] * x + s[3] * y + (r >> 10)) & 63
    c = (s[4] * x + s[5] * y + (r >> 6)) & 63 if count >= 3 else 0
    if a >= b and a >= c: return 0
    if b >= c: return 1
    return 2


def astc(seed, count):
    return [select_partition(seed, i & 3, i >> 2, count) for i in range(16)]


def rows(text):
    return [[int(ch) for ch in row] for row in text.split()]


# BC7 partition tables (2 and 3 subsets), texel order, one row per partition id
BC7_2 = rows("""
0011001100110011 0001000100010001 0111011101110111 0001001100110111 0000000100010011 0011011101111111 0001001101111111 0000000100110111
0000000000010011 0011011111111111 0000000101111111 0000000000010111 0001011111111111 0000000011111111 0000111111111111 0000000000001111
0000100011101111 0111000100000000 0000000010001110 0111001100010000 0011000100000000 0000100011001110 0000000010001100 0111001100110001
0011000100010000 0000100010001100 0110011001100110 0011011001101100 0001011111101000 0000111111110000 0111000110001110 0011100110011100
0101010101010101 0000111100001111 0101101001011010 0011001111001100 0011110000111100 0101010110101010 0110100101101001 0101101010100101
0111001111001110 0001001111001000 0011001001001100 0011101111011100 0110100110010110 0011110011000011 0110011010011001 0000011001100000
0100111001000000 0010011100100000 0000001001110010 0000010011100100 0110110010010011 0011011011001001 0110001110011100 0011100111000110
0110110011001001 0110001100111001 0111111010000001 0001100011100111 0000111100110011 0011001111110000 0010001011101110 0100010001110111
""")
BC7_3 = rows("""
0011001102212222 0001001122112221 0000200122112211 0222002200110111 0000000011221122 0011001100220022 0022002211111111 0011001122112211
0000000011112222 0000111111112222 0000111122222222 0012001200120012 0112011201120112 0122012201220122 0011011211221222 0011200122002220
0001001101121122 0111001120012200 0000112211221122 0022002200221111 0111011102220222 0001000122212221 0000001101220122 0000110022102210
0122012200110000 0012001211222222 0110122112210110 0000011012211221 0022110211020022 0110011020022222 0011012201220011 0000200022112221
0000000211221222 0222002200120011 0011001200220222 0120012001200120 0000111122220000 0120120120120120 0120201212010120 0011220011220011
0011112222000011 0101010122222222 0000000021212121 0022112200221122 0022001100220011 0220122102201221 0101222222220101 0000212121212121
0101010101012222 0222011102220111 0002111200021112 0000211221122112 0222011101110222 0002111211120002 0110011001102222 0000000021122112
0110011022222222 0022001100110022 0022112211220022 0000000000002112 0002000100020001 0222122202221222 0101222222222222 0111201122012220
""")
COMMON2 = [(0, 28, 0), (1, 20, 0), (2, 16, 1), (3, 29, 0), (4, 91, 1), (5, 9, 0), (6, 107, 1), (7, 72, 1), (8, 149, 0), (9, 204, 1),
           (10, 50, 0), (11, 114, 1), (12, 496, 1), (13, 17, 1), (14, 78, 0), (15, 39, 1), (17, 252, 1), (18, 828, 1), (19, 43, 0),
           (20, 156, 0), (21, 116, 0), (22, 210, 1), (23, 476, 1), (24, 273, 0), (25, 684, 1), (26, 359, 0), (29, 246, 1), (32, 195, 1),
           (33, 694, 1), (52, 524, 1)]
COMMON3 = [(4, 260, 0), (8, 74, 5), (9, 32, 5), (10, 156, 2), (11, 183, 2), (12, 15, 0), (13, 745, 4), (20, 0, 1), (35, 335, 1),
           (36, 902, 5), (57, 254, 0)]
PERM3 = [(0, 1, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0), (0, 2, 1), (1, 0, 2)]      # ASTC subset -> BC7 subset
BC7_3_ASTC2 = [(10, 36, 4), (11, 48, 4), (0, 61, 3), (2, 137, 4), (8, 161, 5), (13, 183, 4), (1, 226, 2), (33, 281, 2), (40, 302, 3),
               (20, 307, 4), (21, 479, 0), (58, 495, 3), (3, 593, 0), (32, 594, 2), (59, 605, 1), (34, 799, 3), (20, 812, 1), (14, 988, 4),
               (31, 993, 3)]
MERGE = [(0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 0), (0, 1, 0), (1, 0, 1)]      # BC7 subset -> ASTC subset, by merge id
ANCHORS2 = [(0, 2), (0, 3), (1, 0), (0, 3), (7, 0), (0, 2), (3, 0), (7, 0), (0, 11), (2, 0), (0, 7), (11, 0), (3, 0), (8, 0), (0, 4), (12, 0),
            (1, 0), (8, 0), (0, 1), (0, 2), (0, 4), (8, 0), (1, 0), (0, 2), (4, 0), (0, 1), (4, 0), (1, 0), (4, 0), (1, 0)]


# BC7 anchor ("fix-up") texels of the second subset of the 64 two-subset partitions and of the second / third subset of the 64
# three-subset partitions (the first subset's anchor is always texel 0).  check() verifies that every anchor lies in its subset.
BC7_ANCHOR2 = [15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 2, 8, 2, 2, 8, 8, 15, 2, 8, 2, 2, 8, 8, 2, 2,
               15, 15, 6, 8, 2, 8, 15, 15, 2, 8, 2, 2, 2, 15, 15, 6, 6, 2, 6, 8, 15, 15, 2, 2, 15, 15, 15, 15, 15, 2, 2, 15]
BC7_ANCHOR3A = [3, 3, 15, 15, 8, 3, 15, 15, 8, 8, 6, 6, 6, 5, 3, 3, 3, 3, 8, 15, 3, 3, 6, 10, 5, 8, 8, 6, 8, 5, 15, 15,
                8, 15, 3, 5, 6, 10, 8, 15, 15, 3, 15, 5, 15, 15, 15, 15, 3, 15, 5, 5, 5, 8, 5, 10, 5, 10, 8, 13, 15, 12, 3, 3]
BC7_ANCHOR3B = [15, 8, 8, 3, 15, 15, 3, 8, 15, 15, 15, 15, 15, 15, 15, 8, 15, 8, 15, 3, 15, 8, 15, 8, 3, 15, 6, 10, 15, 15, 10, 8,
                15, 3, 15, 10, 10, 8, 9, 10, 6, 15, 8, 15, 3, 6, 6, 8, 15, 3, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 3, 15, 15, 8]
BC7_W = {2: [0, 21, 43, 64], 3: [0, 9, 18, 27, 37, 46, 55, 64], 4: [0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64]}
ASTC_W = {1: [0, 64], 2: [0, 21, 43, 64], 3: [0, 9, 18, 27, 37, 46, 55, 64], 4: [0, 4, 8, 12, 17, 21, 25, 29, 35, 39, 43, 47, 52, 56, 60, 64],
          5: [0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64]}


def expand7(q):
    return (q << 1) | (q >> 6)


def solid_mode5():
    """For every 8-bit value the pair of 7-bit BC7 mode-5 endpoints that, at colour index 1 (weight 21), reproduces it exactly."""
    out = []
    for v in range(256):
        best = None
        for lo in range(128):
            for hi in range(128):
                got = ((64 - 21) * expand7(lo) + 21 * expand7(hi) + 32) >> 6
                e = abs(got - v)
                if best is None or e < best[0] or (e == best[0] and abs(lo - hi) < abs(best[1] - best[2])):
                    best = (e, lo, hi)
        assert best[0] == 0, v
        out.append((best[1], best[2]))
    return out


def check():
    """The redundancy checks that pin the tables (also run by tests/test_oracle_uastc.py)."""
    for b, s, inv in COMMON2:
        a = astc(s, 2)
        assert [v ^ inv for v in a] == BC7_2[b], (b, s, inv)
    all2 = {tuple(astc(s, 2)) for s in range(1024)}
    found = {b for b in range(64) if tuple(BC7_2[b]) in all2 or tuple(1 - v for v in BC7_2[b]) in all2}
    assert sorted(found) == [b for b, _, _ in COMMON2]
    for b, s, k in COMMON3:
        assert [PERM3[k][v] for v in astc(s, 3)] == BC7_3[b], (b, s, k)
    all3 = {tuple(astc(s, 3)) for s in range(1024)}
    found = {b for b in range(64) for p in itertools.permutations(range(3)) if tuple(p[v] for v in BC7_3[b]) in all3}
    assert sorted(found) == [b for b, _, _ in COMMON3]
    for b, s, k in BC7_3_ASTC2:
        assert [MERGE[k][v] for v in BC7_3[b]] == astc(s, 2), (b, s, k)
    for (b, s, inv), anc in zip(COMMON2, ANCHORS2):
        a = astc(s, 2)
        assert (a.index(0), a.index(1)) == anc, (b, s)
    for b in range(64):          # BC7 anchors lie in the subset they anchor, and texel 0 is always in subset 0
        assert BC7_2[b][0] == 0 and BC7_2[b][BC7_ANCHOR2[b]] == 1, b
        assert BC7_3[b][0] == 0 and BC7_3[b][BC7_ANCHOR3A[b]] == 1 and BC7_3[b][BC7_ANCHOR3B[b]] == 2, b
    return True


def astc_trits(T):
    """The five trits an 8-bit ASTC trit block T encodes (ASTC spec, integer sequence encoding)."""
    b = lambda i: (T >> i) & 1
    if (T >> 2) & 7 == 7:
        C = ((T >> 5) << 2) | (T & 3); t4 = t3 = 2
    else:
        C = T & 31
        if (T >> 5) & 3 == 3: t4 = 2; t3 = b(7)
        else: t4 = b(7); t3 = (T >> 5) & 3
    c = lambda i: (C >> i) & 1
    if C & 3 == 3: t2 = 2; t1 = c(4); t0 = (c(3) << 1) | (c(2) & ~c(3) & 1)
    elif (C >> 2) & 3 == 3: t2 = 2; t1 = 2; t0 = C & 3
    else: t2 = c(4); t1 = (C >> 2) & 3; t0 = (c(1) << 1) | (c(0) & ~c(1) & 1)
    return (t0, t1, t2, t3, t4)


def astc_quints(Q):
    """The three quints a 7-bit ASTC quint block Q encodes."""
    b = lambda i: (Q >> i) & 1
    if (Q >> 1) & 3 == 3 and (Q >> 5) & 3 == 0:
        q2 = (b(0) << 2) | ((b(4) & ~b(0) & 1) << 1) | (b(3) & ~b(0) & 1); q1 = q0 = 4
    else:
        if (Q >> 1) & 3 == 3: q2 = 4; C = (((Q >> 3) & 3) << 3) | ((~(Q >> 5) & 3) << 1) | b(0)
        else: q2 = (Q >> 5) & 3; C = Q & 31
        if C & 7 == 5: q1 = 4; q0 = (C >> 3) & 3
        else: q1 = (C >> 3) & 3; q0 = C & 7
    return (q0, q1, q2)


def bise_encode_tables():
    """Smallest block value per combination (index = plain base-3 / base-5 number, first value = lowest digit).  Checked: every
    combination is reachable with all digits in range, and a bundle whose trailing values are absent (digit 0) has zero in the
    bits a truncated bundle does not store (2 / 4 / 5 / 7 bits for 1..4 trits, 3 / 5 bits for 1 / 2 quints)."""
    trit, quint = {}, {}
    for T in range(256):
        t = astc_trits(T); assert all(0 <= v <= 2 for v in t), (T, t)
        trit.setdefault(sum(v * 3 ** i for i, v in enumerate(t)), T)
    for Q in range(128):
        q = astc_quints(Q); assert all(0 <= v <= 4 for v in q), (Q, q)
        quint.setdefault(sum(v * 5 ** i for i, v in enumerate(q)), Q)
    assert sorted(trit) == list(range(243)) and sorted(quint) == list(range(125))
    for n, bits in ((1, 2), (2, 4), (3, 5), (4, 7)):
        for v in range(3 ** n): assert trit[v] < (1 << bits), (n, v)
    for n, bits in ((1, 3), (2, 5)):
        for v in range(5 ** n): assert quint[v] < (1 << bits), (n, v)
    return [trit[v] for v in range(243)], [quint[v] for v in range(125)]


# ---- ETC2 EAC alpha (UVOL_TEX_ETC2_RGBA: the alpha slice of an ETC1S file re-expressed as an EAC block)
ETC1_INTEN = [(2, 8), (5, 17), (9, 29), (13, 42), (18, 60), (24, 80), (33, 106), (47, 183)]          # (a, b): a block's values are g + {-b, -a, a, b}
EAC_MOD = [[-3, -6, -9, -15, 2, 5, 8, 14], [-3, -7, -10, -13, 2, 6, 9, 12], [-2, -5, -8, -13, 1, 4, 7, 12], [-2, -4, -6, -13, 1, 3, 5, 12],
           [-3, -6, -8, -12, 2, 5, 7, 11], [-3, -7, -9, -11, 2, 6, 8, 10], [-4, -7, -8, -11, 3, 6, 7, 10], [-3, -5, -8, -11, 2, 4, 7, 10],
           [-2, -6, -8, -10, 1, 5, 7, 9], [-2, -5, -8, -10, 1, 4, 7, 9], [-2, -4, -8, -10, 1, 3, 7, 9], [-2, -5, -7, -10, 1, 4, 6, 9],
           [-3, -4, -7, -10, 2, 3, 6, 9], [-1, -2, -3, -10, 0, 1, 2, 9], [-4, -6, -8, -9, 3, 5, 7, 8], [-3, -5, -7, -9, 2, 4, 6, 8]]


def eac_map():
    """For every ETC1S intensity table t and every non-empty set of used selectors (mask): the EAC {table, multiplier, base offset, index
    per selector} with the least squared error on the unclamped values g + {-b, -a, a, b}.  Entry: table | mult << 4 | (offset + 128) << 8 |
    idx0 << 16 | idx1 << 19 | idx2 << 22 | idx3 << 25; rows of EAC_MOD are checked for the format's symmetry (m[4 + k] == -m[k] - 1)."""
    for row in EAC_MOD:
        assert all(row[4 + k] == -row[k] - 1 for k in range(4)), row
    out = []
    for a, b in ETC1_INTEN:
        vals = [-b, -a, a, b]
        for mask in range(16):
            if mask == 0:
                out.append(0); continue
            best = None
            used = [k for k in range(4) if (mask >> k) & 1]
            for tab in range(16):
                for mult in range(1, 16):
                    for off in range(-16, 17):
                        err, idx = 0, [0, 0, 0, 0]
                        for k in used:
                            e, j = min((abs(vals[k] - (off + mult * m)), j) for j, m in enumerate(EAC_MOD[tab]))
                            err += e * e; idx[k] = j
                        key = (err, mult, abs(off), tab)
                        if best is None or key < best[0]:
                            best = (key, tab, mult, off, idx)
            _, tab, mult, off, idx = best
            out.append(tab | (mult << 4) | ((off + 128) << 8) | (idx[0] << 16) | (idx[1] << 19) | (idx[2] << 22) | (idx[3] << 25))
    return out


def unquant_endpoint(val, bits, trits, quints):
    lo, D = val & ((1 << bits) - 1), val >> bits
    if not trits and not quints:
        v = lo << (8 - bits); r = v; sh = bits
        while sh < 8:
            r |= v >> sh; sh += bits
        return r & 255
    A = 511 if lo & 1 else 0
    x = lo >> 1
    if trits:
        C, B = {1: (204, 0), 2: (93, x * 0x116), 3: (44, (x << 7) | (x << 2) | x), 4: (22, (x << 6) | x), 5: (11, (x << 5) | (x >> 2)),
                6: (5, (x << 4) | (x >> 4))}[bits]
    else:
        C, B = {1: (113, 0), 2: (54, x * 0x10C), 3: (26, (x << 7) | (x << 1) | (x >> 1)), 4: (13, (x << 6) | (x >> 1)),
                5: (6, (x << 5) | (x >> 3))}[bits]
    T = (D * C + B) ^ A
    return (A & 0x80) | (T >> 2)


RANGES = [(7, 2, 1, 0), (8, 4, 0, 0), (11, 5, 0, 0), (12, 3, 0, 1), (13, 4, 1, 0), (18, 5, 0, 1), (19, 6, 1, 0), (20, 8, 0, 0)]   # range, bits, trits, quints


def main():
    check()
    pats = [astc(s, 2) for _, s, _ in COMMON2] + [astc(s, 3) for _, s, _ in COMMON3] + [astc(s, 2) for _, s, _ in BC7_3_ASTC2]
    out = ["// uastc_tables.h -- GENERATED by tools/gen/gen_uastc_tables.py (do not edit): UASTC common-partition patterns",
           "// (2 bits per texel, texel i at bits 2i; [0,30) two subsets, [30,41) three subsets, [41,60) mode 7), the anchor mask of",
           "// each pattern (bit i: texel i is the first texel of its subset) and the ASTC endpoint unquantisation of the eight BISE",
           "// ranges UASTC uses (row r: ranges 7, 8, 11, 12, 13, 18, 19, 20; index = low bits | trit-or-quint << bits).",
           "#pragma once", "#include <stdint.h>", "#define UASTC_PAT3_BASE 30", "#define UASTC_PAT7_BASE 41",
           "static const uint32_t UASTC_PATTERN_INIT[60] = {"]
    words = [sum(v << (2 * i) for i, v in enumerate(p)) for p in pats]
    out += ["    " + ", ".join("0x%08xu" % w for w in words[i:i + 6]) + "," for i in range(0, 60, 6)]
    out += ["};", "static const uint16_t UASTC_ANCHOR_INIT[60] = {"]
    masks = [sum(1 << p.index(s) for s in set(p)) for p in pats]
    out += ["    " + ", ".join("0x%04x" % w for w in masks[i:i + 10]) + "," for i in range(0, 60, 10)]
    out += ["};", "static const uint8_t UASTC_UNQUANT_INIT[8][256] = {"]
    for rng, bits, tr, qu in RANGES:
        levels = (3 if tr else 5 if qu else 1) << bits
        row = [unquant_endpoint(v, bits, tr, qu) if v < levels else 0 for v in range(256)]
        out.append("    {" + ", ".join(str(v) for v in row) + "},   // range %d" % rng)
    out.append("};")
    # ---- BC7 target (bc7_core.h): per UASTC pattern the BC7 partition id, how UASTC subsets map to BC7 subsets and the BC7 anchors
    out += ["// BC7 view of the 60 patterns: bits 0-5 BC7 partition id, 6-8 transform (two subsets: invert; three: index into PERM3; mode 7:",
            "// index into MERGE), 9-12 anchor texel of BC7 subset 1, 13-16 anchor texel of BC7 subset 2, 17-22 UASTC subset (2 bits each) that BC7",
            "// subset 0 / 1 / 2 takes its endpoints from.  UASTC_BC7_PAT3: the BC7 three-subset pattern of the mode-7 entries (2 bits per texel).",
            "static const uint32_t UASTC_BC7_INFO_INIT[60] = {"]
    info = []
    for b, s_, inv in COMMON2:
        src = [0 ^ inv, 1 ^ inv, 0]          # BC7 subset x = astc ^ inv  ->  astc = x ^ inv
        info.append(b | (inv << 6) | (BC7_ANCHOR2[b] << 9) | (0 << 13) | (src[0] << 17) | (src[1] << 19) | (src[2] << 21))
    for b, s_, k in COMMON3:
        inv_perm = [PERM3[k].index(x) for x in range(3)]
        info.append(b | (k << 6) | (BC7_ANCHOR3A[b] << 9) | (BC7_ANCHOR3B[b] << 13) | (inv_perm[0] << 17) | (inv_perm[1] << 19) | (inv_perm[2] << 21))
    for b, s_, k in BC7_3_ASTC2:
        info.append(b | (k << 6) | (BC7_ANCHOR3A[b] << 9) | (BC7_ANCHOR3B[b] << 13) | (MERGE[k][0] << 17) | (MERGE[k][1] << 19) | (MERGE[k][2] << 21))
    out += ["    " + ", ".join("0x%08xu" % w for w in info[i:i + 6]) + "," for i in range(0, 60, 6)]
    out += ["};", "static const uint32_t UASTC_BC7_PAT3_INIT[19] = {"]
    p3 = [sum(v << (2 * i) for i, v in enumerate(BC7_3[b])) for b, _, _ in BC7_3_ASTC2]
    out += ["    " + ", ".join("0x%08xu" % w for w in p3[i:i + 6]) + "," for i in range(0, 19, 6)]
    out += ["};", "// ASTC weight index (row = UASTC weight bits 1..5) -> nearest BC7 index of 2 / 3 / 4 index bits", "static const uint8_t UASTC_BC7_WMAP_INIT[3][6][32] = {"]
    for ib in (2, 3, 4):
        out.append("  {")
        for wb in range(6):
            row = [0] * 32
            if wb in ASTC_W:
                for i, wv in enumerate(ASTC_W[wb]):
                    row[i] = min(range(len(BC7_W[ib])), key=lambda j: (abs(BC7_W[ib][j] - wv), j))
            out.append("    {" + ", ".join(str(v) for v in row) + "},")
        out.append("  },")
    out += ["};", "// BC7 mode 5, solid colours: 7-bit endpoints {lo, hi} that give exactly v at colour index 1 (weight 21), v = 0..255", "static const uint8_t BC7_SOLID5_INIT[256][2] = {"]
    sol = solid_mode5()
    out += ["    " + ", ".join("{%d, %d}" % sol[v] for v in range(i, i + 8)) + "," for i in range(0, 256, 8)]
    out.append("};")
    # ---- ASTC target (astc_core.h): the ASTC partition seed of every pattern and the BISE trit / quint block encodings
    out += ["// ASTC view: 10-bit partition seed of the 60 patterns; smallest 8-bit trit block per base-3 number of five trits (first value =",
            "// lowest digit) and smallest 7-bit quint block per base-5 number of three quints (ASTC integer sequence encoding, inverted).",
            "static const uint16_t UASTC_ASTC_SEED_INIT[60] = {"]
    seeds = [s_ for _, s_, _ in COMMON2] + [s_ for _, s_, _ in COMMON3] + [s_ for _, s_, _ in BC7_3_ASTC2]
    out += ["    " + ", ".join(str(v) for v in seeds[i:i + 15]) + "," for i in range(0, 60, 15)]
    tr, qu = bise_encode_tables()
    out += ["};", "static const uint8_t ASTC_TRIT_ENC_INIT[243] = {"]
    out += ["    " + ", ".join(str(v) for v in tr[i:i + 27]) + "," for i in range(0, 243, 27)]
    out += ["};", "static const uint8_t ASTC_QUINT_ENC_INIT[125] = {"]
    out += ["    " + ", ".join(str(v) for v in qu[i:i + 25]) + "," for i in range(0, 125, 25)]
    out.append("};")
    # ---- ETC2 RGBA target (basis_core.h etc1s_alpha_to_eac)
    out += ["// EAC alpha: modifier tables of the format, and per ETC1S intensity table (row) and used-selector mask (column) the EAC block",
            "// parameters that fit the slice's alpha values best: table | mult << 4 | (base offset + 128) << 8 | index of selector k << (16 + 3k).",
            "static const int8_t EAC_MOD_INIT[16][8] = {"]
    out += ["    {" + ", ".join(str(v) for v in row) + "}," for row in EAC_MOD]
    out += ["};", "static const uint32_t ETC1S_EAC_MAP_INIT[8 * 16] = {"]
    em = eac_map()
    out += ["    " + ", ".join("0x%08xu" % w for w in em[i:i + 8]) + "," for i in range(0, 128, 8)]
    out.append("};")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "universal-volumetric_b200", "csrc", "uastc_tables.h")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
