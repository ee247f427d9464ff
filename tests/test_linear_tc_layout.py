"""CPU model of csrc/linear_tc.cu: the index arithmetic of the staging code, the split planes and
the shared-memory descriptors, transcribed expression by expression and executed in numpy.

It cannot prove the kernels right (only a B200 can), but it catches the slips a kernel written
without a GPU is most likely to contain: a wrong plane / stage / K-step offset, a transposed
operand, a thread mapping that leaves holes in a tile. The operand fetch of `tcgen05.mma` is
modelled with the descriptor semantics the tcgen05 probe kernel validated on hardware for this
library's no-swizzle core-matrix layout (csrc/tc_common.cuh):

    K-major  operand, element (row r, k):  start + (r // 8) * SBO + (k // 8) * LBO + (r % 8) * 16 + (k % 8) * 2
    MN-major operand, element (k, col m):  start + (k // 8) * LBO + (m // 8) * SBO + (k % 8) * 16 + (m % 8) * 2
"""

import numpy as np
import pytest

K_CORE = 128


def bf16_bits(a):
    """float32 -> nearest-even bfloat16 bit patterns (uint16)."""
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((b + 0x7FFF + ((b >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def bf16_to_f32(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32)


def split8(v):
    """split8() of the kernel: 8 float32 -> three 16-byte groups of bf16 bit patterns."""
    v = np.asarray(v, dtype=np.float32)
    hi = bf16_bits(v)
    r1 = v - bf16_to_f32(hi)
    mid = bf16_bits(r1)
    r2 = r1 - bf16_to_f32(mid)
    return hi, mid, bf16_bits(r2)


def st_chunk(buf, base, r, cc, C, vals16):
    """tc::st_chunk: 8 bf16 values (16 bytes) of row r, column chunk cc of a tile with C columns."""
    off = base + (r >> 3) * (C >> 3) * K_CORE + cc * K_CORE + (r & 7) * 16
    buf[off // 2: off // 2 + 8] = vals16


def split4(v):
    """split_store4() of the kernel: 4 float32 -> three 8-byte groups of bf16 bit patterns."""
    return split8(v)


def st4(buf, at, vals16):
    """8 bytes (4 bf16 values) at byte offset `at`."""
    buf[at // 2: at // 2 + 4] = vals16


def fetch(buf, start, lbo, sbo, mn_major, rows, k_lo=0):
    """What one K = 16 MMA reads: (rows x 16) float32 through a descriptor (buf is uint16-addressed)."""
    r, k = np.meshgrid(np.arange(rows), np.arange(16), indexing="ij")
    if mn_major:
        a = start + (k // 8) * lbo + (r // 8) * sbo + (k % 8) * 16 + (r % 8) * 2
    else:
        a = start + (r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2
    return bf16_to_f32(buf[a // 2])


PA, PB = (0, 0, 1, 1, 0, 2), (0, 1, 0, 1, 2, 0)


# ------------------------------------------------------------------------------------------
# forward / input-gradient kernel
# ------------------------------------------------------------------------------------------
ROWS, COLS, CHUNK, THREADS = 128, 256, 32, 256
A_TILE, B_TILE = ROWS * CHUNK * 2, COLS * CHUNK * 2
STAGE = 3 * A_TILE + 3 * B_TILE


def prep(w, n_out, k_in, transpose):
    """k_linear_prep."""
    k_chunks, n_tiles = -(-k_in // CHUNK), -(-n_out // COLS)
    planes = np.zeros(n_tiles * k_chunks * 3 * B_TILE // 2, np.uint16)
    groups = k_chunks * 4
    flat = np.asarray(w, np.float32).ravel()
    for i in range(n_tiles * COLS * groups):
        row, grp = i // groups, i % groups
        v = np.zeros(8, np.float32)
        for j in range(8):
            k = grp * 8 + j
            if row < n_out and k < k_in:
                v[j] = flat[k * n_out + row] if transpose else flat[row * k_in + k]
        hi, mid, lo = split8(v)
        tile, r, chunk, cc = row // COLS, row % COLS, grp >> 2, grp & 3
        base = (tile * k_chunks + chunk) * 3 * B_TILE
        st_chunk(planes, base, r, cc, CHUNK, hi)
        st_chunk(planes, base + B_TILE, r, cc, CHUNK, mid)
        st_chunk(planes, base + 2 * B_TILE, r, cc, CHUNK, lo)
    return planes


def linear_cta(x, mask, planes, bias, M, n_out, k_in, act, bx, by, y, x2=None, k_split=None):
    """One CTA of k_linear_tc."""
    k_chunks = -(-k_in // CHUNK)
    k_split = k_in if x2 is None else k_split
    row0, n0 = bx * ROWS, by * COLS
    n_cols = min(COLS, -(-n_out // 16) * 16 - n0)
    b_src = by * k_chunks * 3 * B_TILE
    smem = np.full(2 * STAGE // 2, 0x7FC0, np.uint16)      # NaN-filled: reading an unwritten byte shows up
    acc = np.zeros((ROWS, n_cols), np.float64)
    for c in range(k_chunks):
        s = c & 1
        stage = s * STAGE
        for tid in range(THREADS):
            warp, lane = tid >> 5, tid & 31
            piece = (lane & 1) | ((lane >> 4) << 1)
            kq = ((warp & 1) << 4) | (piece << 2)                # first of the thread's 4 columns in a chunk
            fr0 = ((warp >> 1) << 3) | ((lane >> 1) & 7)         # its first row in the tile
            a_off = (fr0 >> 3) * (CHUNK >> 3) * K_CORE + (kq >> 3) * K_CORE + (fr0 & 7) * 16 + (kq & 7) * 2
            for it in range(4):
                row, k0 = row0 + fr0 + 32 * it, c * CHUNK + kq
                v = np.zeros(4, np.float32)
                if row < M:
                    second = k0 >= k_split
                    left = (k_in if second else k_split) - k0
                    for j in range(4):
                        if j < left:
                            val = x2[row, k0 - k_split + j] if second else x[row, k0 + j]
                            v[j] = 0.0 if (mask is not None and not mask[row, k0 + j] > 0) else val
                hi, mid, lo = split4(v)
                for pl, vals in enumerate((hi, mid, lo)):
                    st4(smem, stage + pl * A_TILE + a_off + it * 2048, vals)
        per_plane = n_cols * 4
        for p in range(3):
            src = (b_src + c * 3 * B_TILE + p * B_TILE) // 2
            dst = (stage + 3 * A_TILE + p * B_TILE) // 2
            smem[dst: dst + per_plane * 8] = planes[src: src + per_plane * 8]
        a0, b0 = stage, stage + 3 * A_TILE
        for k in range(CHUNK // 16):
            for t in range(6):
                a = fetch(smem, a0 + PA[t] * A_TILE + k * 2 * K_CORE, K_CORE, (CHUNK >> 3) * K_CORE, False, ROWS)
                b = fetch(smem, b0 + PB[t] * B_TILE + k * 2 * K_CORE, K_CORE, (CHUNK >> 3) * K_CORE, False, n_cols)
                acc += a.astype(np.float64) @ b.astype(np.float64).T
    for tid in range(THREADS):
        warp = tid >> 5
        r = (warp & 3) * 32 + (tid & 31)
        row, c_lo = row0 + r, (warp >> 2) * 128
        for cb in range(0, 128, 16):
            col = c_lo + cb
            if col >= n_cols:
                break
            if row < M:
                for j in range(16):
                    n = n0 + col + j
                    if n < n_out:
                        o = acc[r, col + j] + (bias[n] if bias is not None else 0.0)
                        y[row, n] = max(o, 0.0) if act else o


@pytest.mark.parametrize("m,k_in,n_out,transpose,masked", [
    (130, 76, 20, False, False),     # K tail (76 = 2 chunks + 12), ragged rows, one narrow column tile
    (40, 40, 260, False, False),     # second column tile (fc9: 256 + 4 outputs)
    (33, 24, 36, True, True),        # input-gradient form: planes of W^T, ReLU mask on the staged operand
])
def test_forward_kernel_model(m, k_in, n_out, transpose, masked):
    rng = np.random.default_rng(m)
    x = rng.standard_normal((m, k_in)).astype(np.float32)
    w = rng.standard_normal((k_in, n_out) if transpose else (n_out, k_in)).astype(np.float32)
    bias = rng.standard_normal(n_out).astype(np.float32)
    mask = rng.standard_normal((m, k_in)).astype(np.float32) if masked else None
    planes = prep(w, n_out, k_in, transpose)
    y = np.full((m, n_out), np.nan)
    for bx in range(-(-m // ROWS)):
        for by in range(-(-n_out // COLS)):
            linear_cta(x, mask, planes, bias, m, n_out, k_in, True, bx, by, y)
    xe = x if mask is None else np.where(mask > 0, x, 0)
    b = w.T if transpose else w
    want = np.maximum(xe.astype(np.float64) @ b.astype(np.float64).T + bias, 0)
    assert not np.isnan(y).any()
    assert np.abs(y - want).max() <= 3e-7 * np.abs(want).max()


def test_forward_kernel_model_two_input_blocks():
    """[x | x2] read in place (fc6: 256 + 76 columns; here 40 + 20 with a ragged tail in each block)."""
    rng = np.random.default_rng(9)
    m, k1, k2, n_out = 20, 40, 20, 24
    x1 = rng.standard_normal((m, k1)).astype(np.float32)
    x2 = rng.standard_normal((m, k2)).astype(np.float32)
    w = rng.standard_normal((n_out, k1 + k2)).astype(np.float32)
    planes = prep(w, n_out, k1 + k2, False)
    y = np.full((m, n_out), np.nan)
    linear_cta(x1, None, planes, None, m, n_out, k1 + k2, False, 0, 0, y, x2=x2, k_split=k1)
    want = np.concatenate([x1, x2], 1).astype(np.float64) @ w.astype(np.float64).T
    assert np.abs(y - want).max() <= 3e-7 * np.abs(want).max()


# ------------------------------------------------------------------------------------------
# weight-gradient kernel
# ------------------------------------------------------------------------------------------
DW_CHUNK, WIDE = 32, 256
DW_TILE = DW_CHUNK * WIDE * 2
DW_STAGE = 6 * DW_TILE


DW_THREADS = 512


def stage_rows(smem, tile, src, mask, row_lo, M, c0, cols, src2=None, split=None, colsum=None):
    for tid in range(DW_THREADS):
        warp, lane = tid >> 5, tid & 31
        ccol, r8 = (warp << 4) | (((lane & 1) | ((lane >> 4) << 1)) << 2), (lane >> 1) & 7
        t_off = (ccol >> 3) * K_CORE + r8 * 16 + (ccol & 7) * 2
        k0 = c0 + ccol
        for a in range(4):
            row = row_lo + r8 + 8 * a
            v = np.zeros(4, np.float32)
            if row < M and k0 < cols:
                sp = cols if src2 is None else split
                second = k0 >= sp
                left = (cols if second else sp) - k0
                for j in range(4):
                    if j < left:
                        val = src2[row, k0 - sp + j] if second else src[row, k0 + j]
                        v[j] = 0.0 if (mask is not None and not mask[row, k0 + j] > 0) else val
            if colsum is not None:
                colsum[tid] += v
            hi, mid, lo = split4(v)
            for pl, vals in enumerate((hi, mid, lo)):
                st4(smem, tile + pl * DW_TILE + t_off + a * 4096, vals)


def dw_cta(dy, mask, x, M, n_out, k_in, bx, gx, by, bz, dw, x2=None, k_split=None, db=None):
    chunks = -(-M // DW_CHUNK)
    per = -(-chunks // gx)
    c_lo, c_hi = bx * per, min(chunks, bx * per + per)
    if c_lo >= c_hi:
        return
    n0, k0 = by * WIDE, bz * WIDE
    m_halves = (min(n_out - n0, WIDE) + 127) // 128
    n_cols = min(WIDE, -(-(k_in - k0) // 16) * 16)
    smem = np.full(2 * DW_STAGE // 2, 0x7FC0, np.uint16)
    acc = np.zeros((2, 128, n_cols), np.float64)
    colsum = np.zeros((DW_THREADS, 4)) if (db is not None and bz == 0) else None
    for it, c in enumerate(range(c_lo, c_hi)):
        stage = (it & 1) * DW_STAGE
        stage_rows(smem, stage, dy, mask, c * DW_CHUNK, M, n0, n_out, colsum=colsum)
        stage_rows(smem, stage + 3 * DW_TILE, x, None, c * DW_CHUNK, M, k0, k_in, x2, k_split)
        a0, b0 = stage, stage + 3 * DW_TILE
        for h in range(m_halves):
            for k in range(DW_CHUNK // 16):
                for t in range(6):
                    koff = k * 2 * (WIDE // 8) * K_CORE
                    a = fetch(smem, a0 + PA[t] * DW_TILE + koff + h * 16 * K_CORE, (WIDE >> 3) * K_CORE, K_CORE, True, 128)
                    b = fetch(smem, b0 + PB[t] * DW_TILE + koff, (WIDE >> 3) * K_CORE, K_CORE, True, n_cols)
                    acc[h] += a.astype(np.float64) @ b.astype(np.float64).T
    if colsum is not None:
        for tid in range(DW_THREADS):
            warp, lane = tid >> 5, tid & 31
            if lane & 14:
                continue
            kdy = n0 + ((warp << 4) | (((lane & 1) | ((lane >> 4) << 1)) << 2))
            total = sum(colsum[tid ^ (q << 1)] for q in range(8))   # the three xor-shuffles (2, 4, 8)
            for j in range(4):
                if kdy + j < n_out:
                    db[kdy + j] += total[j]
    for h in range(m_halves):
        for tid in range(DW_THREADS):
            warp = tid >> 5
            n = n0 + h * 128 + (warp & 3) * 32 + (tid & 31)
            col_lo = (warp >> 2) * 64
            for cb in range(0, 64, 16):
                col = col_lo + cb
                if col >= n_cols:
                    break
                if n < n_out:
                    for j in range(16):
                        if k0 + col + j < k_in:
                            dw[n, k0 + col + j] += acc[h, (warp & 3) * 32 + (tid & 31), col + j]


@pytest.mark.parametrize("m,n_out,k_in,gx,masked", [
    (70, 20, 40, 2, True),       # three 32-row chunks over two slabs (one CTA gets two chunks), ragged tail
    (33, 132, 24, 4, False),     # second M = 128 accumulator; more slabs than chunks (empty CTAs)
    (20, 260, 270, 1, False),    # second block along n_out (grid.y) and along k_in (grid.z)
])
def test_weight_gradient_kernel_model(m, n_out, k_in, gx, masked):
    rng = np.random.default_rng(n_out)
    dy = rng.standard_normal((m, n_out)).astype(np.float32)
    x = rng.standard_normal((m, k_in)).astype(np.float32)
    mask = rng.standard_normal((m, n_out)).astype(np.float32) if masked else None
    dw, db = np.zeros((n_out, k_in)), np.zeros(n_out)
    for bx in range(gx):
        for by in range(-(-n_out // WIDE)):
            for bz in range(-(-k_in // WIDE)):
                dw_cta(dy, mask, x, m, n_out, k_in, bx, gx, by, bz, dw, db=db)
    dye = dy if mask is None else np.where(mask > 0, dy, 0)
    want = dye.astype(np.float64).T @ x.astype(np.float64)
    assert np.abs(dw - want).max() <= 3e-7 * np.abs(want).max()
    assert np.abs(db - dye.astype(np.float64).sum(0)).max() <= 1e-6 * np.abs(dye).sum(0).max()


def test_weight_gradient_kernel_model_two_input_blocks():
    """X = [x | x2] with the split inside the second 256-column block (fc6: 256 + 76)."""
    rng = np.random.default_rng(4)
    m, n_out, k1, k2 = 40, 12, 264, 20
    dy = rng.standard_normal((m, n_out)).astype(np.float32)
    x1 = rng.standard_normal((m, k1)).astype(np.float32)
    x2 = rng.standard_normal((m, k2)).astype(np.float32)
    dw = np.zeros((n_out, k1 + k2))
    for bz in range(2):
        dw_cta(dy, None, x1, m, n_out, k1 + k2, 0, 1, 0, bz, dw, x2=x2, k_split=k1)
    want = dy.astype(np.float64).T @ np.concatenate([x1, x2], 1).astype(np.float64)
    assert np.abs(dw - want).max() <= 3e-7 * np.abs(want).max()
