"""Arithmetic facts the CUDA kernels rely on, checked on the CPU by enumeration (no GPU, no library call).

Each test names the place in csrc/ that leans on the fact; the GPU parity tests would catch a violation only for the inputs they
happen to draw."""
import numpy as np

from oracle import preprocess as opre

# every source size the reference's resize rule can meet in practice, plus odd ones
SIZES = [(1280, 720), (1920, 1080), (3840, 2160), (640, 360), (854, 480), (1920, 800), (512, 288), (320, 180), (1000, 562), (130, 70),
         (720, 1280), (1080, 1920), (257, 145), (4096, 2160), (2560, 1440), (960, 540), (848, 480), (1440, 1080)]


def _all_weight_pairs():
    pairs = set()
    for w, h in SIZES:
        nw, nh = opre.target_size(w, h)
        for src, dst, clamp in ((w, nw, True), (h, nh, False)):
            if src == dst or src == 2 * dst:
                continue
            _, _, w0, w1 = opre.linear_coeffs(src, dst, clamp)
            pairs.update(zip(w0.tolist(), w1.tolist()))
    return sorted(pairs)


def test_tap_weights_are_non_negative_and_sum_to_2048():
    """OpenCV's 11-bit coefficients of a tap pair (cvRound of (1 - f) * 2048 and f * 2048, f in [0, 1)) add up to 2048 for every
    geometry above -- and to 2048 +- 1 for ANY float32 fraction (f = (m + 0.5) / 2048 + one ulp rounds both weights up: 2049).
    preprocess.cu (quad kernel, `<= 255 by construction`) and conv_tc.cu (bilinear_word) drop the saturating cast; the next test
    shows that this holds up to a sum of 2049 in both passes."""
    pairs = _all_weight_pairs()
    assert len(pairs) > 100
    for a, b in pairs:
        assert a >= 0 and b >= 0 and a + b == 2048, (a, b)
    f = np.arange(0, 2048 * 64, dtype=np.float64) / (2048 * 64)
    f32 = f.astype(np.float32)
    f = np.concatenate([f32, np.nextafter(f32, np.float32(1)), np.nextafter(f32, np.float32(0))])
    f = f[(f >= 0) & (f < 1)]
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    w1 = np.rint(f * np.float32(2048)).astype(np.int64)
    assert (w0 >= 0).all() and (w1 >= 0).all()
    assert (np.abs(w0 + w1 - 2048) <= 1).all()
    assert (w0 + w1 == 2049).any()                       # the pathological fractions exist: the bound below must cover them


def test_vertical_pass_needs_no_clamp_and_the_rounding_folds_into_the_product():
    """out = (((b0 (S0 >> 4)) >> 16) + ((b1 (S1 >> 4)) >> 16) + 2) >> 2 with S = a0 p + a1 q, bytes p, q, and weight pairs that sum
    to 2049 at most: (i) 0 <= out <= 255, so no clamp; (ii) ((b0 u0 + (2 << 16)) >> 16) == ((b0 u0) >> 16) + 2, the form the
    kernels compute (the constant rides in the multiply-add); (iii) everything stays inside int32."""
    s_max = 2049 * 255                                    # a0 + a1 <= 2049, both bytes 255
    u_max = s_max >> 4
    u = np.unique(np.concatenate([np.arange(0, 4096), np.arange(u_max - 4096, u_max + 1),
                                  np.random.default_rng(0).integers(0, u_max + 1, 20000)])).astype(np.int64)
    for total in (2047, 2048, 2049):
        for b0 in list(range(0, total + 1, 7)) + [1, total - 1, total, total // 2, total // 2 + 1]:
            b1 = total - b0
            p0 = b0 * u
            assert p0.max() + (2 << 16) < 2 ** 31
            assert (((p0 + (2 << 16)) >> 16) == (p0 >> 16) + 2).all()
            # worst case: both rows at their extremes
            hi = (((b0 * u_max) >> 16) + ((b1 * u_max) >> 16) + 2) >> 2
            assert 0 <= hi <= 255, (total, b0, hi)
    # the 2x2 mean of the quad kernel: (sum of four bytes + 2) >> 2 <= 255
    assert (4 * 255 + 2) >> 2 == 255


def test_operand_ring_wrap_and_capacity():
    """conv_tc.cu FrRing<768>: positions wrap by x - 768 * umulhi(x, MAGIC) (768 is not a power of two), exact for every position a
    launch can reach; and 768 positions keep the ring deadlock-free with all eight unfold warps busy for every output width the
    fused kernel accepts (a tile's views span 128 + 2 P1w positions; the unfold may overwrite a position only after the last tile
    that reads it, and that tile must not itself need the row being written)."""
    cap = 768
    magic = (1 << 32) // cap + 1
    x = np.arange(0, 1 << 22, dtype=np.uint64)
    assert ((x - cap * ((x * magic) >> 32)) == x % cap).all()
    for dst_w in range(3, 257):
        p1w = (dst_w + 2) // 3
        assert cap > 3 * p1w + 126                                   # the tile a writer waits for never needs the writer's row
        assert cap >= 3 * p1w + 126 + -(-8 * p1w // 3)               # ... with eight rows (8/3 pooled rows) in the warps' hands
    # the quad kernel's unaligned staging: whole chunks from the aligned address below a row cover it, with room for the zero chunk
    for row_bytes in (390, 2562, 3000, 3840, 5760, 11520):
        row_pad = ((row_bytes + 15) // 16 + 1) * 16
        quad_pad = row_pad + 32
        for lead in range(16):
            nch = (lead + row_bytes + 15) >> 4
            assert 16 * (nch + 1) <= quad_pad                        # chunks [0, nch) + one chunk of zeros fit the slot
            assert 16 * nch >= lead + row_bytes                      # ... and cover the row
            assert lead + 3 * (row_bytes // 3 - 1) + 12 <= 16 * (nch + 1)      # the last pixel's three words stay inside
