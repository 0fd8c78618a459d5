"""Pure-Python Philox4x32-10 and the sequential draw stream used to drive the executed reference.
TEST INFRASTRUCTURE ONLY."""
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
DOM_SEQ = 7


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


class SeqStream:
    """draw i = philox(ctr=(i_lo, i_hi, stream_id, DOM_SEQ), key=seed); one block per draw."""

    def __init__(self, seed, stream_id=0):
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.stream_id = stream_id
        self.i = 0

    def _block(self):
        x = philox4x32_10((self.i & MASK, (self.i >> 32) & MASK, self.stream_id, DOM_SEQ), self.key)
        self.i += 1
        return x

    def randbelow(self, n):
        return (self._block()[0] * n) >> 32

    def u53(self):
        x = self._block()
        return ((x[1] >> 5) << 26) | (x[2] >> 6)
