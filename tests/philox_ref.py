"""TEST INFRASTRUCTURE: Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11) in plain Python — the checker of b200kge_sample_uniform (kge_b200/csrc/rowwise.cu)."""
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c = list(counter)
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def sample_uniform(n, K, vocab, seed, offset):
    """Mirror of sample_uniform_kernel: element e uses block (e // 2, offset) under key seed, words (0,1) | (2,3)."""
    out = []
    for e in range(n * K):
        pair = e // 2
        c = philox4x32_10([pair & MASK, (pair >> 32) & MASK, offset & MASK, (offset >> 32) & MASK],
                          (seed & MASK, (seed >> 32) & MASK))
        r = (c[1] << 32 | c[0]) if e % 2 == 0 else (c[3] << 32 | c[2])
        out.append((r * vocab) >> 64)
    return out
