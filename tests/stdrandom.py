"""Bit-exact Python model of std::mt19937_64 + libstdc++'s std::uniform_real_distribution<double>,
the PRNG the reference's unit tests seed with 42 (e.g. src/reconstruction/plmc_cuda_tests.cu:37-38)."""

_M64 = (1 << 64) - 1


class MT19937_64:
    def __init__(self, seed):
        self.mt = [0] * 312
        self.mt[0] = seed & _M64
        for i in range(1, 312):
            self.mt[i] = (6364136223846793005 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 62)) + i) & _M64
        self.idx = 312

    def _twist(self):
        mt = self.mt
        for i in range(312):
            x = (mt[i] & 0xFFFFFFFF80000000) | (mt[(i + 1) % 312] & 0x7FFFFFFF)
            xa = x >> 1
            if x & 1:
                xa ^= 0xB5026F5AA96619E9
            mt[i] = mt[(i + 156) % 312] ^ xa
        self.idx = 0

    def __call__(self):
        if self.idx >= 312:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= (y >> 29) & 0x5555555555555555
        y ^= (y << 17) & 0x71D67FFFEDA60000
        y ^= (y << 37) & 0xFFF7EEE000000000
        y ^= y >> 43
        return y & _M64


class UniformReal:
    """libstdc++: generate_canonical<double,53>(g) = double(g()) / 2^64 (one draw), clamped below 1;
    result = canon * (b - a) + a."""

    def __init__(self, a, b):
        self.a, self.b = a, b

    def __call__(self, gen):
        canon = float(gen()) / 18446744073709551616.0
        if canon >= 1.0:
            canon = 1.0 - 2.0 ** -53
        return canon * (self.b - self.a) + self.a


def mt19937_uniform_real(seed, a, b, count):
    """`count` draws of std::uniform_real_distribution<double>(a, b) on a 32-bit std::mt19937(seed), as libstdc++ evaluates
    them: generate_canonical<double, 53> takes two engine outputs, (x1 + x2 * 2^32) / 2^64 in double arithmetic, clamped
    below 1.  numpy's MT19937 with its legacy seeding is init_genrand(seed), the engine's seeding."""
    import numpy as np
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    raw = bg.random_raw(2 * count).astype(np.uint64)
    canon = (raw[0::2].astype(np.float64) + raw[1::2].astype(np.float64) * 4294967296.0) / 18446744073709551616.0
    canon[canon >= 1.0] = np.nextafter(1.0, 0.0)
    return canon * (b - a) + a
