"""GPU parity tests of the device-side Groestl-256 Merkle commitment against the oracle:
leaf digests for the leaf sizes the prover uses, pair compressions, whole trees, layers and branches
(mirrors crates/core/src/merkle_tree/tests.rs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    import binius_b200

    layer = binius_b200.B200Layer(0)
    yield layer
    layer.close()


def _digests(hal, d, n):
    return [bytes(r) for r in np.ascontiguousarray(hal.to_host(d)).view(np.uint8).reshape(n, 32)]


@pytest.mark.parametrize("leaf_elems", [1, 2, 3, 4, 5, 7, 8, 16, 64])
def test_leaf_digests(hal, oracle, leaf_elems):
    # 16..1024-byte leaves: every padding case of 16-byte-aligned messages (rem = 0, 16, 32, 48)
    n_leaves = 300
    elems = oracle.rand_b128(10 + leaf_elems, n_leaves * leaf_elems)
    d, out = hal.to_device(elems), hal.dev_alloc(2 * n_leaves)
    hal._check(hal._lib.b200_groestl256_leaves(hal._ctx, d.ptr, n_leaves, leaf_elems, out.ptr))
    raw = elems.tobytes()
    got = _digests(hal, out, n_leaves)
    for i in range(n_leaves):
        assert got[i] == oracle.groestl256(raw[16 * leaf_elems * i: 16 * leaf_elems * (i + 1)]), i


def test_pair_compression(hal, oracle):
    n = 1000
    x = oracle.rand_b128(77, 4 * n)
    d, out = hal.to_device(x), hal.dev_alloc(2 * n)
    hal._check(hal._lib.b200_groestl256_compress_pairs(hal._ctx, d.ptr, n, out.ptr))
    raw = x.tobytes()
    got = _digests(hal, out, n)
    for i in range(n):
        assert got[i] == oracle.groestl256_compress_pair(raw[64 * i: 64 * i + 32], raw[64 * i + 32: 64 * i + 64])


@pytest.mark.parametrize("log_len,batch", [(0, 4), (1, 1), (6, 16), (12, 16), (14, 2)])
def test_merkle_tree_matches_oracle(hal, oracle, log_len, batch):
    from binius_b200.merkle import BinaryMerkleTree

    elems = oracle.rand_b128(90 + log_len, batch << log_len)
    tree = BinaryMerkleTree.build(hal, hal.to_device(elems), batch)
    exp = oracle.merkle_build(elems, batch)
    assert tree.root() == exp[-1]
    assert _digests(hal, tree.nodes, len(exp)) == exp
    for depth in {0, log_len // 2, log_len}:
        n_nodes = len(exp)
        start = n_nodes + 1 - (1 << (depth + 1))
        assert tree.layer(depth) == exp[start:start + (1 << depth)]
    # a branch verifies against the root (scheme.rs verify_opening: fold the leaf digest up with the siblings)
    for index in {0, (1 << log_len) - 1, (1 << log_len) // 3}:
        br = tree.branch(index, 0)
        node = exp[index]
        for j, sib in enumerate(br):
            node = oracle.groestl256_compress_pair(node, sib) if ((index >> j) & 1) == 0 else oracle.groestl256_compress_pair(sib, node)
        assert node == exp[-1]


def test_merkle_errors(hal, oracle):
    import binius_b200
    from binius_b200.merkle import BinaryMerkleTree

    d = hal.to_device(oracle.rand_b128(1, 48))
    with pytest.raises(binius_b200.InputValidation):
        BinaryMerkleTree.build(hal, d, 5)  # IncorrectBatchSize
    with pytest.raises(binius_b200.InputValidation):
        BinaryMerkleTree.build(hal, d, 16)  # 3 leaves: PowerOfTwoLengthRequired


def test_codeword_commit_2pow20_leaves(hal, oracle):
    """commit shape of the prover (fri/prove.rs:120-198): 2^24 B128 codeword elements in cosets of 16 -> 2^20 leaves of
    256 bytes; spot-checked leaves + the upper layers recomputed by the oracle from the device's layer 10"""
    from binius_b200.merkle import BinaryMerkleTree

    n, batch = 1 << 24, 16
    dev = hal.dev_alloc(n)
    hal.fill(dev, 0)
    seed = oracle.rand_b128(123, 1 << 16)
    for k in range(0, n, 1 << 16):  # 256 distinct blocks: block k is the seed folded by a different challenge
        blk = dev.slice(k, k + (1 << 16))
        hal.copy_h2d(seed ^ np.uint64(k + 1), blk)
    tree = BinaryMerkleTree.build(hal, dev, batch)
    raw = hal.to_host(dev.slice(0, 1 << 16)).tobytes()
    leaves = _digests(hal, tree.nodes.slice(0, 2 * 4096), 4096)
    for i in (0, 1, 1000, 4095):
        assert leaves[i] == oracle.groestl256(raw[256 * i: 256 * (i + 1)])
    layer10 = tree.layer(10)
    cur = layer10
    while len(cur) > 1:
        cur = [oracle.groestl256_compress_pair(cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)]
    assert cur[0] == tree.root()
