"""Host-side Philox4x32-10 (the device generator's twin) for the few draws the host makes per move: the greedy
tie-break (domain 2) and the agent's start particle (domain 3).  Counter = (block, id, move, domain), key = seed."""
_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
DOM_SIM, DOM_UPDATE, DOM_FINAL, DOM_AGENT = 0, 1, 2, 3


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = (int(c) & _MASK for c in ctr)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def key_of(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & _MASK, seed >> 32


def randbelow(seed, n, block, ident, move, domain):
    """Integer in [0, n): multiply-shift of word 0 of the block."""
    return (philox4x32_10((block, ident, move, domain), key_of(seed))[0] * int(n)) >> 32
