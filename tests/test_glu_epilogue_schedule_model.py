"""CPU model of the work distribution of the gated-activation epilogue (csrc/gemm_tcgen05.cu::glu_epilogue_tile
and the epilogue branch `if constexpr (EPI)` of gemm_kernel / gemm2_kernel).

The CUDA kernel itself is checked bit for bit on the B200 (tests/test_gpu_gemm_lora.py); what this test pins on a
CPU-only box is the SCHEDULE the code implements -- which warp touches which 16-column unit of which tile, and
that the two-register-buffer operand prefetch (incl. the hand-off of the NEXT tile's first unit during the last
unit of the current tile) always delivers the operands of exactly the unit that is computed next:

  * every element (row < M, col < N) is produced exactly once, rows >= M never;
  * every `finish(buffer, row, col)` consumes a buffer whose most recent prefetch was for that same (row, col);
  * the persistent tile walk (tile = w % num_tiles, stride = number of CTAs / CTA pairs) covers every tile once.

The loops below are a line-by-line transcription of the device code with the TMEM / global accesses replaced by
bookkeeping."""
import itertools

import pytest

GLU_WARPS_PER_QUARTER = 4
BLOCK_M = 128


def epilogue_tile(block_n, q0, row, row_ok, col_tile, sub, nrow, ncol, log):
    """glu_epilogue_tile<BLOCK_N>: q0 / q1 are dicts {'row', 'col'} standing for the register buffers."""
    units = block_n // (16 * GLU_WARPS_PER_QUARTER)

    def prefetch(q, r, c):
        q["row"], q["col"] = r, c

    def finish(q, r, c):
        assert (q.get("row"), q.get("col")) == (r, c), ("stale operands", q, r, c)
        log.append((r, c))

    if units == 1:
        if row_ok:
            finish(q0, row, col_tile + 16 * sub)
        if nrow >= 0:
            prefetch(q0, nrow, ncol)
        return
    q1 = {}
    for j in range(0, units, 2):
        cA = 16 * sub + 64 * j
        cB = cA + 64
        if row_ok:
            prefetch(q1, row, col_tile + cB)
        if row_ok:
            finish(q0, row, col_tile + cA)
        if j + 2 < units:
            if row_ok:
                prefetch(q0, row, col_tile + cB + 64)
        elif nrow >= 0:
            prefetch(q0, nrow, ncol)
        if row_ok:
            finish(q1, row, col_tile + cB)


def run_kernel(M, N, block_n, pair, n_ctas):
    """Every epilogue thread of every CTA of the persistent launch; returns the list of produced (row, col0) units."""
    tile_m = 2 * BLOCK_M if pair else BLOCK_M
    m_tiles = (M + tile_m - 1) // tile_m
    n_tiles = N // block_n
    num_work = m_tiles * n_tiles
    walkers = min(num_work, n_ctas // 2 if pair else n_ctas)
    produced = []
    for walker in range(walkers):
        for rank in range(2 if pair else 1):
            for q, sub, lane in itertools.product(range(4), range(GLU_WARPS_PER_QUARTER), (0, 13, 31)):
                def coords(w):
                    tile = w % num_work
                    m_idx, n_idx = tile % m_tiles, tile // m_tiles       # any bijection serves the model
                    return m_idx * tile_m + rank * BLOCK_M + q * 32 + lane, n_idx * block_n
                q0, log = {}, []
                if walker < num_work:
                    row, col_tile = coords(walker)
                    if row < M:
                        q0.update(row=row, col=col_tile + 16 * sub)
                w = walker
                while w < num_work:
                    row, col_tile = coords(w)
                    nrow, ncol = -1, 0
                    if w + walkers < num_work:
                        nr, nc = coords(w + walkers)
                        if nr < M:
                            nrow, ncol = nr, nc + 16 * sub
                    epilogue_tile(block_n, q0, row, row < M, col_tile, sub, nrow, ncol, log)
                    w += walkers
                produced += log
    return produced


@pytest.mark.parametrize("M,N,block_n,pair,n_ctas", [
    (8192, 14336, 256, True, 148), (300, 512, 256, True, 148), (1000, 1024, 256, True, 148),
    (640, 384, 128, True, 148), (304, 256, 256, False, 148), (136, 128, 128, False, 148),
    (136, 64, 64, False, 148), (5000, 768, 256, True, 6), (520, 512, 128, False, 3), (129, 192, 64, False, 2)])
def test_every_unit_is_produced_once_from_fresh_operands(M, N, block_n, pair, n_ctas):
    if M * N > 4_000_000:                       # the cfg2 shape: sample the schedule on a narrower matrix
        M, N = 1536, 1024 if block_n == 256 else N
    produced = run_kernel(M, N, block_n, pair, n_ctas)
    lanes = (0, 13, 31)
    want = {(r, c) for r in range(M) if (r % 32) in lanes for c in range(0, N, 16)}
    assert len(produced) == len(set(produced)), "a unit was produced twice"
    assert set(produced) == want
