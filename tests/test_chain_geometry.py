"""Host logic of the persistent decoder-step kernel: the stream-K partition (csrc/dec_chain.cu::chain_geom).

Every (weight tile, k-block) unit must belong to exactly one CTA, ranges must be contiguous, and the number of CTAs that
contribute a partial tile to one output tile must never exceed the slots reserved for it -- the consumers add the slots in
contributor order, so a wrong count would silently drop or double a partial sum.  Runs without a GPU.
"""
import ctypes as C

import pytest

import open_whisper_kit_b200 as pkg

DIMS = [384, 512, 768, 1024, 1280]


@pytest.fixture(scope="module")
def lib():
    return pkg.load(strict_api=False)


def geom(lib, grid, rows, N, K, min_units, direct=0):
    out = (C.c_int * 5)()
    rc = lib.whisper_b200_chain_geometry(grid, rows, N, K, min_units, direct, out)
    return rc, tuple(out)


@pytest.mark.parametrize("d", DIMS)
@pytest.mark.parametrize("grid", [148, 296, 264])
@pytest.mark.parametrize("min_units", [1, 2, 4, 10])
def test_stream_k_partition_covers_every_unit_once(lib, d, grid, min_units):
    for N, K in ((d, d), (d, 4 * d), (3 * d, d), (4 * d, d)):
        rc, (tiles, kpt, U, G, maxc) = geom(lib, grid, 64, N, K, min_units)
        assert rc == 0
        assert tiles == N // 128 and kpt == K // 64 and U == tiles * kpt
        assert 1 <= G <= min(grid, max(1, U))
        start = [U * c // G for c in range(G + 1)]          # range_of() in the kernel
        assert start[0] == 0 and start[-1] == U and all(b >= a for a, b in zip(start, start[1:]))
        if U >= G:
            assert all(b > a for a, b in zip(start, start[1:]))       # nobody idles inside the participating set
        owner = [None] * U
        for c in range(G):
            for u in range(start[c], start[c + 1]):
                assert owner[u] is None
                owner[u] = c
        assert all(o is not None for o in owner)
        for u in range(U):                                   # sg_cta_of() in the kernel
            assert ((u + 1) * G - 1) // U == owner[u]
        for t in range(tiles):
            contributors = {owner[u] for u in range(t * kpt, (t + 1) * kpt)}
            first, last = owner[t * kpt], owner[(t + 1) * kpt - 1]
            assert contributors == set(range(first, last + 1))        # contiguous, so slot = cta - first
            assert len(contributors) <= maxc
        assert (U + 1) * G < 2 ** 31                         # the kernel does this arithmetic in 32 bits


@pytest.mark.parametrize("d", DIMS)
def test_direct_geometry_one_tile_per_cta(lib, d):
    rc, (tiles, kpt, U, G, maxc) = geom(lib, 296, 64, 3 * d, d, 4, direct=1)
    assert rc == 0 and tiles == 3 * d // 128 and G == tiles and U == tiles * kpt and maxc == 0
    assert all((U * (c + 1) // G) - (U * c // G) == kpt for c in range(G))


def test_rejects_shapes_the_kernel_does_not_take(lib):
    assert geom(lib, 296, 64, 1280 + 64, 1280, 4)[0] == -1       # N not a multiple of the 128-row weight tile
    assert geom(lib, 296, 64, 1280, 1280 + 32, 4)[0] == -1       # K not a multiple of 64
    assert geom(lib, 296, 129, 1280, 1280, 4)[0] == -1           # more rows than one accumulator holds
    assert geom(lib, 0, 64, 1280, 1280, 4)[0] == -1
