"""Host-side model of the index arithmetic of the LayerNorm warps inside gemm_tc2_kernel<..., LNW> (mint_b200/csrc/
gemm_tc.cu): which warp normalises which rows of a 256-row block, which arrival counters it waits for, and who clears
them.  The kernel itself is tested on the GPU (tests/test_kernels_gpu.py::test_gemm_layernorm_inside_the_launch); this
checks, for every shape the host code can send there, the properties a mistake in that arithmetic would break silently
or by deadlock: every row below M is normalised exactly once, every counter that is waited on does get its arrivals,
and every counter is cleared exactly once, by the last unit past its wait."""
BM2, LN_WARPS, EPI_HALVES = 256, 8, 2


def _units(tiles_n):
    return tiles_n * 2 * LN_WARPS


def _unit_of_row(r, units):
    return ((r + 1) * units + BM2 - 1) // BM2 - 1


def _simulate(M, tiles_n):
    units = _units(tiles_n)
    groups = (M + 31) // 32
    arrivals = [0] * (groups + 1)
    # epilogue warps: two per 32-row group and column tile, if the group starts below M
    for block in range((M + BM2 - 1) // BM2):
        for g in range(8):
            if block * BM2 + g * 32 < M:
                arrivals[(block * BM2 + g * 32) // 32] += EPI_HALVES * tiles_n
    rows_done = [0] * M
    retired = [0] * (groups + 1)
    cleared = [0] * (groups + 1)
    for block in range((M + BM2 - 1) // BM2):
        base = block * BM2
        for u in range(units):                      # u = (tile % tiles_n) * 16 + rank * 8 + warp
            r_lo, r_hi = u * BM2 // units, (u + 1) * BM2 // units
            if r_hi == r_lo or base + (r_lo & ~31) >= M:
                continue
            g_lo, g_hi = r_lo >> 5, (r_hi - 1) >> 5
            waited = [g for g in range(g_lo, g_hi + 1) if base + g * 32 < M]
            for g in waited:
                gi = (base + g * 32) // 32
                assert arrivals[gi] == EPI_HALVES * tiles_n, "waits for a counter nobody fills"
                waiters = _unit_of_row(g * 32 + 31, units) - _unit_of_row(g * 32, units) + 1
                retired[gi] += 1
                if retired[gi] == waiters:
                    cleared[gi] += 1
            for r in range(base + r_lo, min(base + r_hi, M)):
                rows_done[r] += 1
    assert all(c == 1 for c in rows_done)
    for gi in range(groups):
        assert cleared[gi] == 1, (M, tiles_n, gi, retired[gi])
    assert cleared[groups] == 0


def test_every_row_once_and_every_counter_cleared_once():
    for tiles_n in (1, 2, 3, 4, 5, 6):              # n = 160 .. 960 (the host only takes n % 160 == 0, n <= 1024)
        assert _units(tiles_n) <= BM2
        for M in (1, 31, 32, 33, 127, 128, 129, 255, 256, 257, 360, 3840, 7680, 11520, 19999, 20000, 46080):
            _simulate(M, tiles_n)


def test_unit_of_row_inverts_the_row_ranges():
    for tiles_n in (1, 3, 5, 6):
        units = _units(tiles_n)
        for u in range(units):
            for r in range(u * BM2 // units, (u + 1) * BM2 // units):
                assert _unit_of_row(r, units) == u
