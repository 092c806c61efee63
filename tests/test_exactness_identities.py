"""Float32 identities the kernels lean on where they do not evaluate the reference's expression literally
(DESIGN.md section 3); checked here on the CPU in IEEE binary32 with gradual underflow, which is what the CUDA
build uses (-ftz=false, round-to-nearest intrinsics)."""
import numpy as np


def _finite_bit_patterns(rng, n):
    bits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    return x[np.isfinite(x)]


def test_sign_of_a_rounded_difference_is_the_order_of_its_operands():
    """k_tri's footprint decides coverage by comparing the row term with the column term (cov_test): for finite a, b,
    fl(a - b) < 0 exactly when a < b -- also next to each other, in the denormal range, across signs and at +-0."""
    rng = np.random.default_rng(11)
    a = _finite_bit_patterns(rng, 1 << 21)
    b = _finite_bit_patterns(rng, 1 << 21)
    n = min(len(a), len(b))
    a, b = a[:n], b[:n]
    neighbours = np.nextafter(a, np.float32(np.inf), dtype=np.float32)
    tiny = (rng.integers(0, 1 << 24, n, dtype=np.uint64).astype(np.uint32)).view(np.float32)   # zero and denormals
    zeros = np.array([0.0, -0.0, 1.0, -1.0, 1e-45, -1e-45], np.float32)
    with np.errstate(over="ignore"):   # a - b may overflow to +-inf: the sign is still the order
        for x, y in [(a, b), (a, neighbours), (neighbours, a), (a, a), (tiny, tiny[::-1]), (a, -a),
                     (np.repeat(zeros, len(zeros)), np.tile(zeros, len(zeros)))]:
            ok = np.isfinite(x) & np.isfinite(y)
            x, y = x[ok], y[ok]
            d = (x - y).astype(np.float32)
            assert d.dtype == np.float32
            assert np.array_equal(d < 0, x < y)
            assert np.array_equal(~(d < 0), x >= y)   # "not negative" == ">=": what setp.geu tests for finite operands


def test_candidate_coordinates_by_float_addition():
    """(float)(minx + k) == (float)minx + (float)k for every column / row index a frame can have (sizes are u16) and the
    footprint's k = 0..2: both sides are integers below 2^24, so the float addition is exact."""
    m = np.arange(0, 1 << 16, dtype=np.uint32)
    for k in range(3):
        assert np.array_equal((m + k).astype(np.float32), m.astype(np.float32) + np.float32(k))
