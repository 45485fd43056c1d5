"""The Groestl-256 oracle (oracle/groestl.c, written from the specification) against the published known answers of
Groestl-256, and the structure of the Merkle restatement (binary_merkle_tree.rs:27-164).  The reference itself holds no
stored Groestl vectors (groestl/tests.rs compares with the groestl crate on random inputs)."""


def test_groestl256_known_answers(oracle):
    # Groestl-256 KATs of the SHA-3 submission package (ShortMsgKAT_256: Len = 0; and the "abc" test vector)
    assert oracle.groestl256(b"").hex() == "1a52d11d550039be16107f9c58db9ebcc417f16f736adb2502567119f0083467"
    assert oracle.groestl256(b"abc").hex() == "f3c1bb19c048801326a7efbcf16e3d7887446249829c379e1840d1a3a1e7d4d2"


def test_padding_boundaries_are_consistent(oracle):
    # 55 / 56 / 63 / 64 / 65 bytes cross the one-vs-two padding block boundary: all digests distinct, deterministic
    seen = set()
    for n in (1, 55, 56, 63, 64, 65, 119, 120, 128, 256):
        d = oracle.groestl256(bytes(range(256))[:n] if n <= 256 else b"")
        assert d == oracle.groestl256(bytes(range(256))[:n])
        seen.add(d)
    assert len(seen) == 10


def test_merkle_tree_structure(oracle):
    elems = oracle.rand_b128(5, 64)
    nodes = oracle.merkle_build(elems, 4)  # 16 leaves of 4 elements
    assert len(nodes) == 31
    raw = elems.tobytes()
    for i in range(16):
        assert nodes[i] == oracle.groestl256(raw[64 * i: 64 * (i + 1)])
    off, n = 0, 16
    while n > 1:
        for i in range(n // 2):
            assert nodes[off + n + i] == oracle.groestl256_compress_pair(nodes[off + 2 * i], nodes[off + 2 * i + 1])
        off, n = off + n, n // 2
