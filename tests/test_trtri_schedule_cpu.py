"""The schedule of the blocked triangular inversion (bopy_b200/csrc/fit_kernels.cuh, tile_gemm_async_kernel modes 1 / 2, and
launch_trtri in bopy_b200.cu) restated in numpy: levels h = 1, 2, 4, ...; per level every (pair, row-in-bottom-half,
column-in-top-half) tile, its k range, and the scratch T parked in the strict upper block triangle of W.  Guards the index
arithmetic for block counts that are not powers of two (the kernels themselves are checked on the GPU by the LML-gradient and
inverse-path tests)."""
import numpy as np
import pytest


def blocked_inverse_like_the_kernels(L, bs):
    n = L.shape[0]
    nb = n // bs
    W = np.zeros_like(L)

    def blk(M, i, j):
        return M[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]

    for i in range(nb):                                    # copy_diag_blocks_kernel: W_II = inv(L_II)
        blk(W, i, i)[:] = np.linalg.inv(blk(L, i, i))
    h = 1
    while h < nb:
        grid = ((nb + 2 * h - 1) // (2 * h)) * h * h
        for mode in (1, 2):                                # two launches per level; tiles of a launch are independent
            new = {}
            for b in range(grid):
                pair, rem = divmod(b, h * h)
                top, bottom = 2 * h * pair, 2 * h * pair + h
                i, j = bottom + rem // h, top + rem % h
                if i >= nb:
                    continue
                if mode == 1:                              # T_IJ = sum_{K=J}^{bottom-1} L_IK W_KJ  -> parked at block (J, I)
                    acc = sum(blk(L, i, k) @ blk(W, k, j) for k in range(j, bottom))
                    new[(j, i)] = acc
                else:                                      # W_IJ = -sum_{K=bottom}^{I} W_IK T_KJ, T_KJ read from block (J, K)
                    acc = sum(blk(W, i, k) @ blk(W, j, k) for k in range(bottom, i + 1))
                    new[(i, j)] = -acc
            for (r, c), v in new.items():                  # (stores of one launch never feed loads of the same launch)
                blk(W, r, c)[:] = v
        h *= 2
    return W


@pytest.mark.parametrize("nb", list(range(1, 12)) + [16, 21, 64])
def test_recursive_block_inversion_schedule(nb):
    bs = 3
    rng = np.random.default_rng(nb)
    n = nb * bs
    L = np.tril(rng.standard_normal((n, n))) + 4.0 * np.eye(n)
    W = blocked_inverse_like_the_kernels(L, bs)
    np.testing.assert_allclose(np.tril(W), np.linalg.inv(L), rtol=1e-9, atol=1e-11)
    for i in range(nb):                                    # diagonal blocks keep a clean upper triangle; the scratch lives
        d = W[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs]    # in the strict upper BLOCK triangle only
        assert np.array_equal(np.triu(d, 1), np.zeros_like(d))


@pytest.mark.parametrize("n,grid", [(33, 3), (50, 4), (256, 16), (333, 21), (700, 44), (2048, 128), (3000, 148), (8192, 148)])
def test_probe_inv_row_deal_covers_every_row_once_and_is_balanced(n, grid):
    """probe_inv_kernel.cuh: rows of W = L^-1 are dealt to the 16 * grid warps boustrophedon; row r costs r + 1 products."""
    warps = 16 * grid
    assert grid == min(148, (n + 15) // 16)              # inv_grid() in bopy_b200.cu
    owner = np.full(n, -1)
    load = np.zeros(warps, dtype=np.int64)
    for gw in range(warps):
        k = 0
        while k * warps < n:
            r = (k + 1) * warps - 1 - gw if k & 1 else k * warps + gw
            if r < n:
                assert owner[r] == -1
                owner[r] = gw
                load[gw] += r + 1
            k += 1
    assert (owner >= 0).all()
    if n >= 4 * warps:                                    # several full sweeps: every warp gets about the same number of products
        assert load.max() <= 1.2 * load.mean()
