"""CPU: integer scan-path tables of the product package, bit exact against the reference's."""
import hashlib

import numpy as np
import pytest

import zigma_b200
from util import gold


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 8, 16, 32, 64])
def test_zigzag_bit_exact(N):
    g = gold("tables")
    mine = np.stack(zigma_b200.zigzag_path(N))
    assert mine.dtype == np.int64
    assert np.array_equal(mine, g[f"zigzag_{N}"])
    for p in mine:   # permutation + inverse round trip
        rev = zigma_b200.reverse_permut_np(p)
        assert np.array_equal(p[rev], np.arange(N * N)) and np.array_equal(rev[p], np.arange(N * N))


@pytest.mark.parametrize("N", [2, 3, 4, 6, 8, 16, 32])
def test_hilbert_bit_exact(N):
    g = gold("tables")
    mine = np.stack(zigma_b200.hilbert_path(N))
    assert np.array_equal(mine, g[f"hilbert_{N}"])
    for p in mine:
        assert np.array_equal(np.sort(p), np.arange(N * N))


def test_zigzag32_sha256():
    z = np.stack(zigma_b200.zigzag_path(32)).astype(np.int64)
    assert hashlib.sha256(z.tobytes()).hexdigest().startswith("01b6ef874ac9cd89")
    g = gold("tables")
    assert bytes(g["sha256_zigzag_32_int64"]) == hashlib.sha256(z.tobytes()).digest()


def test_zigzag_is_continuous():
    """Every consecutive pair of a zigzag path is a 4-neighbour step (the point of the ZigMa paths)."""
    N = 16
    for p in zigma_b200.zigzag_path(N):
        r, c = p // N, p % N
        assert np.all(np.abs(np.diff(r)) + np.abs(np.diff(c)) == 1)
