"""Host-side restatement of the opaque saved-gates layout of the bf16 GRU kernels (include/fadernets_b200.h,
FnGruChainBf16.gates; `gate_off` in csrc/fn_gru_tc.cu): [32 rows][16 columns] blocks, block order (time slab, 32-row block,
16-column block).  Checks that the map is a bijection onto T * ceil32(B) * 4H elements, that a time segment starts at
t0 * ceil32(B) * 4H (the offset the decoder wavefront passes), and that a warp of the gate epilogue (32 consecutive rows x
16 consecutive units) lands in ONE contiguous 1 KB block."""
import numpy as np
import pytest


def gate_off(t, b, col, B, H4):
    row_blocks = (B + 31) >> 5
    return ((t * row_blocks + (b >> 5)) * (H4 >> 4) + (col >> 4)) * 512 + (b & 31) * 16 + (col & 15)


@pytest.mark.parametrize("T,B,H", [(3, 64, 32), (2, 70, 64), (4, 5, 16), (1, 256, 64)])
def test_gate_layout_is_a_bijection(T, B, H):
    H4, Bp = 4 * H, (B + 31) // 32 * 32
    t, b, c = np.meshgrid(np.arange(T), np.arange(Bp), np.arange(H4), indexing="ij")
    off = gate_off(t, b, c, B, H4).ravel()
    assert off.min() == 0 and off.max() == T * Bp * H4 - 1
    assert np.unique(off).size == off.size
    # a time segment [t0, ...) starts at t0 * ceil32(B) * 4H
    for t0 in range(T):
        assert gate_off(t0, 0, 0, B, H4) == t0 * Bp * H4


def test_epilogue_warp_access_is_one_contiguous_block():
    B, H = 256, 1024
    H4 = 4 * H
    for row0 in (0, 32, 224):
        for col0 in (0, 16, H + 48, 3 * H + 1008):
            rows, cols = np.meshgrid(np.arange(row0, row0 + 32), np.arange(col0, col0 + 16), indexing="ij")
            off = np.sort(gate_off(7, rows, cols, B, H4).ravel())
            assert off[0] % 512 == 0 and np.array_equal(off, off[0] + np.arange(512))
