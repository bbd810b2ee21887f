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
