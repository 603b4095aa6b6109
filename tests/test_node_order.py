"""CPU: the node walk of the register variant (glmmadaptive_b200/csrc/kernels.cuh, `node_id`), restated.

The kernel visits the K = nAGQ^q quadrature nodes in an order chosen so that the 16-byte shared-memory loads of a
quarter warp (eight consecutive lanes = eight consecutive positions of the walk) never collide: the rows of the
first dimension's table that the eight lanes read are either the same row or rows whose indices differ mod 8
(row pitch = an odd number of 16-byte units).  This test pins the two properties the kernel relies on -- the walk
is a permutation of the nodes, and every aligned group of eight positions is conflict-free in dimension 0 and
reads at most a handful of distinct rows in the other dimensions -- and the position of the central node (the
round that supplies the log-sum-exp shift)."""
import itertools

import pytest


def node_id(k, nagq, K):
    """position in the walk -> natural node index a_0 + nAGQ (a_1 + nAGQ a_2) (kernels.cuh: node_id)"""
    nfull = nagq & ~7
    nrem = nagq - nfull
    nhi = K // nagq
    if k < nfull * nhi:
        return (k % nfull) + nagq * (k // nfull)
    k2 = k - nfull * nhi
    return nfull + (k2 % nrem) + nagq * (k2 // nrem)


def central_position(nagq, K):
    nfull = nagq & ~7
    nrem = nagq - nfull
    nhi = K // nagq
    c = (K - 1) >> 1
    a0c, hic = c % nagq, c // nagq
    return hic * nfull + a0c if a0c < nfull else nfull * nhi + hic * nrem + (a0c - nfull)


@pytest.mark.parametrize("q,nagq", [(q, n) for q, n in itertools.product((1, 2, 3), range(1, 17)) if n ** q <= 512])
def test_walk_is_a_conflict_free_permutation(q, nagq):
    K = nagq ** q
    ids = [node_id(k, nagq, K) for k in range(K)]
    assert sorted(ids) == list(range(K))
    assert ids[central_position(nagq, K)] == (K - 1) >> 1
    if nagq < 8:
        assert ids == list(range(K))          # fewer than eight nodes per dimension: the natural order
    for g0 in range(0, K, 8):
        group = ids[g0:g0 + 8]
        a0 = [i % nagq for i in group]
        # dimension 0: equal rows are one broadcast address, different rows must differ mod 8
        distinct = sorted(set(a0))
        assert len({a % 8 for a in distinct}) == len(distinct), (g0, a0)
        if q > 1:
            # the other dimensions see consecutive indices: at most ceil(8 / (nAGQ mod 8)) + 1 distinct rows
            a1 = sorted({(i // nagq) % nagq for i in group})
            assert len(a1) <= 9
