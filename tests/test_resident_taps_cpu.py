"""Index arithmetic behind the resident-weight conv kernels (csrc/conv_tc_slab.cuh, ConvSlabCfg RES): in the expanded weight
matrix B[kh][w][co][(w', c)] of k_prep_conv_weights (tap kw = w' - w + PADL) the K-slice of the input column w' = 3 - PADL
holds the taps kw = 3, 2, 1, 0 as consecutive COUT-row blocks, and the operand of every other (w', kh) is the contiguous
sub-range of those blocks that starts at block 3 - w' + wlo(w') - PADL.  NumPy only."""
import numpy as np
import pytest


@pytest.mark.parametrize("PADL", [1, 2])
@pytest.mark.parametrize("KH,CIN,COUT", [(3, 32, 48), (2, 16, 32), (5, 16, 32), (3, 64, 32), (2, 32, 16), (5, 32, 16)])
def test_every_operand_is_a_block_range_of_the_full_column(PADL, KH, CIN, COUT):
    rng = np.random.default_rng(KH * 100 + CIN + PADL)
    NOUT, KROW = 4 * COUT, 4 * CIN
    taps = rng.standard_normal((KH, 4, CIN, COUT)).astype(np.float32)
    B = np.zeros((KH * NOUT, KROW), np.float32)
    for kh in range(KH):
        for w in range(4):
            for wp in range(4):
                kw = wp - w + PADL
                if 0 <= kw <= 3:
                    B[kh * NOUT + w * COUT:kh * NOUT + (w + 1) * COUT, wp * CIN:(wp + 1) * CIN] = taps[kh, kw].T
    full = 3 - PADL
    for kh in range(KH):
        resident = B[kh * NOUT:(kh + 1) * NOUT, full * CIN:(full + 1) * CIN]
        for wp in range(4):
            wl, wh = max(0, wp - (3 - PADL)), min(3, wp + PADL)
            nb = wh - wl + 1
            ring_box = B[kh * NOUT + wl * COUT:kh * NOUT + (wl + nb) * COUT, wp * CIN:(wp + 1) * CIN]
            j0 = 3 - wp + wl - PADL
            assert 0 <= j0 and j0 + nb <= 4
            assert np.array_equal(ring_box, resident[j0 * COUT:(j0 + nb) * COUT])
