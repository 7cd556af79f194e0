"""Host-side model of the shared-memory "slab" layout used by every tensor-core operand
(see csrc/sm100.cuh): R rows x 64 half columns, 128 B per row, 16-byte chunk c of row r stored at
chunk position c ^ (r % 8), rows grouped by 8 into 1024-byte atoms.  Used to pack weights and by
the bring-up tests; numpy only."""
import numpy as np

SWIZZLE_128B = 2


def smem_desc_template(lbo_bytes, sbo_bytes, swizzle=SWIZZLE_128B):
    return (((lbo_bytes >> 4) & 0x3FFF) << 16) | (((sbo_bytes >> 4) & 0x3FFF) << 32) | (1 << 46) | ((swizzle & 7) << 61)


def idesc_f16(m, n, a_mn_major=0, b_mn_major=0, a_bf16=0, b_bf16=0):
    return ((1 << 4) | (a_bf16 << 7) | (b_bf16 << 10) | (a_mn_major << 15) | (b_mn_major << 16)
            | ((n >> 3) << 17) | ((m >> 4) << 24))


def pack_slab(x):
    """x: (R, 64) float16 (R multiple of 8) -> uint8 image of R*128 bytes."""
    x = np.ascontiguousarray(x, dtype=np.float16)
    r, c = x.shape
    assert c == 64 and r % 8 == 0
    chunks = x.view(np.uint8).reshape(r, 8, 16)           # (row, chunk, 16 B)
    rows = np.arange(r)
    out = np.empty_like(chunks)
    for ch in range(8):
        out[rows, ch ^ (rows & 7)] = chunks[rows, ch]
    return out.reshape(-1)


def unpack_slab(img, rows):
    chunks = np.frombuffer(np.ascontiguousarray(img), dtype=np.uint8).reshape(rows, 8, 16)
    r = np.arange(rows)
    out = np.empty_like(chunks)
    for ch in range(8):
        out[r, ch] = chunks[r, ch ^ (r & 7)]
    return out.reshape(rows, 128).view(np.float16)


def pack_matrix(x):
    """x: (R, K) float16 with K % 64 == 0 -> concatenation of K/64 slabs (slab j = columns 64j..64j+63)."""
    r, k = x.shape
    assert k % 64 == 0
    return np.concatenate([pack_slab(x[:, j:j + 64]) for j in range(0, k, 64)])
