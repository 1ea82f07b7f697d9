"""numpy replica of the kernel's noise stream: Philox4x32-10 keyed by the seed, counter
(k*6 + t//2, 0, pedestrian, 0), Box-Muller on uniforms built from the top 23 bits with the
exponent trick: u1 = 2 - float(0x3f800000 | a >> 9) in (0, 1], u2 = float(0x3f800000 | b >> 9) - 1."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    c = [np.asarray(x, np.uint64) for x in ctr]
    k0, k1 = int(key[0]), int(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def _box_muller(a, b):
    one_two = lambda x: ((x >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)      # [1, 2)
    u1 = np.float32(2.0) - one_two(a)
    u2 = one_two(b) - np.float32(1.0)
    r = np.sqrt(np.float64(-1.3862943611198906) * np.log2(u1.astype(np.float64))).astype(np.float32)
    ang = (np.float32(6.283185307179586) * u2).astype(np.float64)
    return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)


def philox_normals(seed, peds, K):
    """-> (K, 12, len(peds), 2) float32, the eps the kernel draws for those global pedestrian ids."""
    peds = np.asarray(peds, np.uint64)
    out = np.empty((K, 12, peds.size, 2), np.float32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    for k in range(K):
        for j in range(6):
            c0 = np.full(peds.shape, k * 6 + j, np.uint64)
            z = np.zeros_like(peds)
            r = philox4x32_10([c0, z, peds, z], key)
            out[k, 2 * j, :, 0], out[k, 2 * j, :, 1] = _box_muller(r[0], r[1])
            out[k, 2 * j + 1, :, 0], out[k, 2 * j + 1, :, 1] = _box_muller(r[2], r[3])
    return out
