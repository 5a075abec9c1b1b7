"""Host-side decoders of the tiled fp16 hi/lo operand layouts (include/mepp_b200.h) -- used by
tests and debugging only; the kernels produce and consume the tiled form directly."""
import numpy as np


def decode_x_planes(raw_u8, m_rows, n_pad, with_flags=False):
    """[m_tile][k_block][plane 2][kc 4][row 128][8] fp16 (+ the non-zero bit array behind the tiles) ->
    (hi, lo) float32 arrays (m_rows, n_pad); with_flags also returns the bit array as bool (MT, KB, 2)."""
    MT = (m_rows + 127) // 128
    KB = n_pad // 32
    raw = np.ascontiguousarray(raw_u8).view(np.uint8)
    a = raw[:MT * KB * 16384].view(np.float16).reshape(MT, KB, 2, 4, 128, 8).astype(np.float32)
    # -> (plane, MT, row, KB, kc, e)
    a = a.transpose(2, 0, 4, 1, 3, 5).reshape(2, MT * 128, KB * 32)
    if with_flags:
        FW = ((2 * KB + 31) // 32 + 3) // 4 * 4
        words = raw[MT * KB * 16384:MT * KB * 16384 + MT * FW * 4].view(np.uint32).reshape(MT, FW)
        bits = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(MT, FW * 32)
        return a[0][:m_rows], a[1][:m_rows], bits[:, :2 * KB].reshape(MT, KB, 2).astype(bool)
    return a[0][:m_rows], a[1][:m_rows]


def decode_y_planes(raw_u8, C, n_pad, BN, n_tiles_n):
    """[n_tile][k_block][plane 2][kc 4][col BN][8] fp16 -> (hi, lo) float32 arrays (n_pad, C)."""
    KB = n_pad // 32
    a = np.frombuffer(np.ascontiguousarray(raw_u8).tobytes(), dtype=np.float16)[:n_tiles_n * KB * 2 * 4 * BN * 8]
    a = a.reshape(n_tiles_n, KB, 2, 4, BN, 8).astype(np.float32)
    # -> (plane, KB, kc, e, NT, col)
    a = a.transpose(2, 1, 3, 5, 0, 4).reshape(2, KB * 32, n_tiles_n * BN)
    return a[0][:, :C], a[1][:, :C]


def decode_symbols(sym_T, N, L):
    """sym_T uint32 [W3][n_pad] (3 bits per base, continuous stream per sequence) -> int8 (N, L) symbols 0..4."""
    w = np.ascontiguousarray(sym_T).view(np.uint32).astype(np.uint64)
    out = np.zeros((N, L), dtype=np.int8)
    for i in range(L):
        bit = 3 * i
        wi, off = bit >> 5, bit & 31
        v = (w[wi, :N] | (w[wi + 1, :N] << np.uint64(32))) >> np.uint64(off)
        out[:, i] = (v & np.uint64(7)).astype(np.int8)
    return out
