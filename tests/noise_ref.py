"""CPU restatement of itg_noise_normal (TEST INFRASTRUCTURE): Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
as 1, 2, 3", SC'11; Random123 v1.x `philox4x32_R(10, ...)`) + Box-Muller, vectorised with numpy."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
# Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds): (counter, key) -> output
KAT = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
       ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
       ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Arrays (or scalars) of uint64 holding 32-bit values -> four uint64 arrays."""
    c = [np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for r in range(10):
        if r > 0:
            k0, k1 = (k0 + np.uint64(W0)) & MASK, (k1 + np.uint64(W1)) & MASK
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
    return c


def noise_window(shape_full, window, seed, field):
    C, Hf, Wf = shape_full
    y0, y1, x0, x1 = window
    cc, yy, xx = np.meshgrid(np.arange(C, dtype=np.uint64), np.arange(y0, y1, dtype=np.uint64), np.arange(x0, x1, dtype=np.uint64), indexing="ij")
    e = (cc * np.uint64(Hf) + yy) * np.uint64(Wf) + xx
    g = e >> np.uint64(2)
    out = philox4x32_10(g & MASK, g >> np.uint64(32), np.full_like(g, field), np.zeros_like(g), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    pair = (e & np.uint64(2)) != 0
    a = np.where(pair, out[2], out[0]).astype(np.float32)
    b = np.where(pair, out[3], out[1]).astype(np.float32)
    u1 = (a + np.float32(1.0)) * np.float32(2.3283064365386963e-10)
    u2 = b * np.float32(2.3283064365386963e-10)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1.astype(np.float64))).astype(np.float32)
    ang = 2.0 * np.pi * u2.astype(np.float64)
    odd = (e & np.uint64(1)) != 0
    return (rad * np.where(odd, np.sin(ang), np.cos(ang))).astype(np.float32)
