"""CPU: the arithmetic behind the certified bounds of the dk/dn screening pass (more3d_b200/csrc/features.cu,
ScreenAcc), emulated in numpy with the kernel's float32 operation order.  No device needed."""
import numpy as np

F = np.float32


def pack10(n):
    """q = rint(511 c) + 512 per component (feat_kword); returns the three 10-bit fields."""
    return (np.rint(n * 511.0).clip(-511, 511) + 512).astype(np.uint32)


def fma32(a, b, c):
    """float32 fused multiply-add: the float64 product of two float32 is exact, one rounding at the end."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def screened_t(n_p, q):
    """t = max(1 - n_p . n_q', 0) as ScreenAcc::set_normal / add evaluate it (fields read in place as mantissa bits)."""
    kx, ky, kz = 2097152.0, 2048.0, 1024.0
    ax, ay, az = F(n_p[:, 0] * kx / 511.0), F(n_p[:, 1] * ky / 511.0), F(n_p[:, 2] * kz / 511.0)
    c0 = F(1.0 + (n_p[:, 0] * (kx + 512.0) + n_p[:, 1] * (ky + 512.0) + n_p[:, 2] * (kz + 512.0)) / 511.0)
    one = np.uint32(0x3F800000)
    gx = (one | (q[:, 0] << np.uint32(2))).view(F)           # field at bits 2..11: 1 + q 2^-21
    gy = (one | (q[:, 1] << np.uint32(12))).view(F)          # bits 12..21: 1 + q 2^-11
    gz = (one | (q[:, 2] << np.uint32(13))).view(F)          # bits 22..31 shifted down by 9: 1 + q 2^-10
    t = fma32(-gz, az, fma32(-gy, ay, fma32(-gx, ax, c0)))
    return np.maximum(t, F(0.0))


def unit(v):
    return v / np.linalg.norm(v, axis=1)[:, None]


def test_packed_normal_bound():
    """|t - dn| <= 2.3e-3 for dn = 1 - n_p . n_q (the delta of ScreenAcc::dn_settled_by_chi2), on random pairs, on
    near-parallel pairs (the terrain case) and on axis-aligned normals (components at +-1, the quantiser's ends)."""
    rng = np.random.Generator(np.random.PCG64(7))
    m = 400000
    a = unit(rng.normal(size=(m, 3)))
    b = unit(rng.normal(size=(m, 3)))
    near = unit(a + 0.08 * rng.normal(size=(m, 3)))
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float)
    ax_p = axes[rng.integers(0, 6, m)]
    ax_q = unit(ax_p + 1e-3 * rng.normal(size=(m, 3)))
    worst = 0.0
    for p, q in ((a, b), (a, near), (ax_p, ax_q), (ax_p, axes[rng.integers(0, 6, m)])):
        dn = 1.0 - np.einsum("ij,ij->i", p, q)
        t = screened_t(p, pack10(q)).astype(np.float64)
        err = np.abs(t - np.maximum(dn, 0.0))
        worst = max(worst, float(err.max()))
    assert worst <= 2.3e-3, worst
    assert worst > 5e-4          # the bound is not vacuous: quantisation alone reaches ~1.7e-3


def test_std_bound_is_an_upper_bound():
    """std(dn) <= sqrt(var_t (1 + 2 eps) + 4 eps mean_t^2) + 2.3e-3 with float32 sums of t and t^2 (dn_settled_by_chi2)."""
    rng = np.random.Generator(np.random.PCG64(11))
    for trial in range(300):
        mu = int(rng.integers(4, 400))
        spread = 10 ** rng.uniform(-3, -0.3)
        p = unit(rng.normal(size=(1, 3)))
        q = unit(p + spread * rng.normal(size=(mu, 3)))
        dn = 1.0 - q @ p[0]
        t = screened_t(np.repeat(p, mu, 0), pack10(q))
        s1, s2 = F(0), F(0)
        for v in t:                                   # the kernel's sequential float32 accumulation
            s1 = F(s1 + v)
            s2 = F(np.float64(v) * np.float64(v) + np.float64(s2))
        mean = float(s1) / mu
        var = max(float(s2) / mu - mean * mean, 0.0)
        eps = (mu + 2.0) / 8388608.0
        std_up = np.sqrt(var * (1 + 2 * eps) + 4 * eps * mean * mean) + 2.3e-3
        assert dn.std() <= std_up, (trial, dn.std(), std_up)


def test_float32_interval_matches_the_fp64_rule():
    """{K_q float32 : |fl64(K_p - K_q)| <= T} is the closed float32 interval dk_interval finds: brute force over the
    float32 neighbours of both ends."""
    rng = np.random.Generator(np.random.PCG64(13))
    for T in (0.04, np.nextafter(0.0004, 1.0)):
        for Kp in np.concatenate([rng.normal(0, 0.02, 200), [0.0, 0.04, -0.04, 1e-30, 3.0e5]]).astype(F):
            Kp64 = np.float64(Kp)
            pred = lambda x: abs(Kp64 - np.float64(x)) <= T
            lo = F(Kp64 - T)
            while pred(np.nextafter(lo, F(-np.inf))):
                lo = np.nextafter(lo, F(-np.inf))
            while not pred(lo):
                lo = np.nextafter(lo, F(np.inf))
            hi = F(Kp64 + T)
            while pred(np.nextafter(hi, F(np.inf))):
                hi = np.nextafter(hi, F(np.inf))
            while not pred(hi):
                hi = np.nextafter(hi, F(-np.inf))
            xs = np.concatenate([lo + np.arange(-3, 4) * np.spacing(lo), hi + np.arange(-3, 4) * np.spacing(hi),
                                 rng.normal(Kp, 0.05, 50).astype(F)]).astype(F)
            for x in xs:
                assert pred(x) == (lo <= x <= hi), (Kp, T, x, lo, hi)
