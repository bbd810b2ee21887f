"""ETC1 block decoder (Khronos OES_compressed_ETC1_RGB8_texture), numpy, TEST INFRASTRUCTURE.  Used to pin the ETC1 target format:
decoding the library's ETC1 blocks must give exactly the oracle's RGBA32 texels of the same ETC1S source."""
import numpy as np

MOD = np.array([[2, 8, -2, -8], [5, 17, -5, -17], [9, 29, -9, -29], [13, 42, -13, -42], [18, 60, -18, -60], [24, 80, -24, -80],
                [33, 106, -33, -106], [47, 183, -47, -183]], np.int32)             # table codeword -> modifier by pixel index value


def decode_etc1(blocks, width, height):
    """blocks u8[nblocks, 8] in block raster order -> u8[height, width, 4] (alpha 255)."""
    bx, by = (width + 3) // 4, (height + 3) // 4
    assert blocks.shape == (bx * by, 8)
    b = blocks.astype(np.int32)
    diff, flip = (b[:, 3] >> 1) & 1, b[:, 3] & 1
    t1, t2 = (b[:, 3] >> 5) & 7, (b[:, 3] >> 2) & 7
    base = np.zeros((len(b), 2, 3), np.int32)
    for c in range(3):
        hi5, d3 = b[:, c] >> 3, b[:, c] & 7
        d3 = np.where(d3 >= 4, d3 - 8, d3)
        c1d = (hi5 << 3) | (hi5 >> 2); v2 = hi5 + d3; c2d = (v2 << 3) | (v2 >> 2)                     # differential: 5 bits + signed 3-bit delta
        a4, b4 = b[:, c] >> 4, b[:, c] & 15
        c1i, c2i = (a4 << 4) | a4, (b4 << 4) | b4                                                     # individual: 4 + 4 bits
        base[:, 0, c] = np.where(diff == 1, c1d, c1i); base[:, 1, c] = np.where(diff == 1, c2d, c2i)
    msb = (b[:, 4] << 8) | b[:, 5]; lsb = (b[:, 6] << 8) | b[:, 7]
    out = np.zeros((by * 4, bx * 4, 4), np.uint8); out[..., 3] = 255
    for x in range(4):
        for y in range(4):
            p = x * 4 + y
            idx = (((msb >> p) & 1) << 1) | ((lsb >> p) & 1)
            sub = np.where(flip == 1, 1 if y >= 2 else 0, 1 if x >= 2 else 0)                        # flip: top / bottom halves, else left / right
            table = np.where(sub == 1, t2, t1)
            mod = MOD[table, idx]
            col = np.clip(base[np.arange(len(b)), sub] + mod[:, None], 0, 255).astype(np.uint8)
            out[y::4, x::4, :3] = col.reshape(by, bx, 3)
    return out[:height, :width]


EAC_MOD = np.array([[-3, -6, -9, -15, 2, 5, 8, 14], [-3, -7, -10, -13, 2, 6, 9, 12], [-2, -5, -8, -13, 1, 4, 7, 12], [-2, -4, -6, -13, 1, 3, 5, 12],
                    [-3, -6, -8, -12, 2, 5, 7, 11], [-3, -7, -9, -11, 2, 6, 8, 10], [-4, -7, -8, -11, 3, 6, 7, 10], [-3, -5, -8, -11, 2, 4, 7, 10],
                    [-2, -6, -8, -10, 1, 5, 7, 9], [-2, -5, -8, -10, 1, 4, 7, 9], [-2, -4, -8, -10, 1, 3, 7, 9], [-2, -5, -7, -10, 1, 4, 6, 9],
                    [-3, -4, -7, -10, 2, 3, 6, 9], [-1, -2, -3, -10, 0, 1, 2, 9], [-4, -6, -8, -9, 3, 5, 7, 8], [-3, -5, -7, -9, 2, 4, 6, 8]], np.int32)


def decode_etc2_rgba(blocks, width, height):
    """ETC2 RGBA8 (COMPRESSED_RGBA8_ETC2_EAC) blocks u8[nblocks, 16] in block raster order -> u8[height, width, 4]: bytes 0-7 the EAC alpha
    block (base, multiplier | table, 16 x 3-bit indices, pixel x * 4 + y first in the top bits: alpha = clamp(base + modifier * multiplier)),
    bytes 8-15 the colour block, decoded as ETC1 (the differential / individual modes; an ETC2 block in T / H / planar mode is not what
    the product writes and trips the assertion)."""
    bx, by = (width + 3) // 4, (height + 3) // 4
    assert blocks.shape == (bx * by, 16)
    col = blocks[:, 8:].astype(np.int32)
    diff = (col[:, 3] >> 1) & 1
    for c in range(3):          # differential blocks must not overflow 5 bits (that is how ETC2 marks its T / H / planar modes)
        hi5, d3 = col[:, c] >> 3, col[:, c] & 7
        v2 = hi5 + np.where(d3 >= 4, d3 - 8, d3)
        assert ((diff == 0) | ((v2 >= 0) & (v2 <= 31))).all()
    out = decode_etc1(blocks[:, 8:], width, height).copy()
    a = blocks[:, :8].astype(np.int64)
    base, mult, table = a[:, 0], a[:, 1] >> 4, a[:, 1] & 15
    bits = np.zeros(len(a), np.int64)
    for k in range(2, 8):
        bits = (bits << 8) | a[:, k]
    alpha = np.zeros((by * 4, bx * 4), np.uint8)
    for x in range(4):
        for y in range(4):
            idx = (bits >> (45 - 3 * (x * 4 + y))) & 7
            v = np.clip(base + EAC_MOD[table, idx] * mult, 0, 255).astype(np.uint8)
            alpha[y::4, x::4] = v.reshape(by, bx)
    out[..., 3] = alpha[:height, :width]
    return out
