"""Binary Merkle tree over device-resident data with Groestl-256 (reference crates/core/src/merkle_tree/
binary_merkle_tree.rs:16-164, BinaryMerkleTreeProver crates/core/src/merkle_tree/prover.rs): the committed RS codeword
and the FRI round oracles are hashed where they live, only the root / opened branches cross PCIe."""
from __future__ import annotations

from typing import List

import numpy as np

from .layer import B200Layer, DevSlice, InputValidation


class BinaryMerkleTree:
    """`inner_nodes` = leaf layer followed by every inner layer, root last (binary_merkle_tree.rs:21-26);
    a digest occupies 2 B128 slots of the device buffer."""

    def __init__(self, hal: B200Layer, log_len: int, nodes: DevSlice):
        self.hal, self.log_len, self.nodes = hal, log_len, nodes

    @classmethod
    def build(cls, hal: B200Layer, elements: DevSlice, batch_size: int) -> "BinaryMerkleTree":
        """BinaryMerkleTree::build (binary_merkle_tree.rs:27-56)"""
        if batch_size <= 0 or elements.len() % batch_size:
            raise InputValidation("IncorrectBatchSize")
        n_leaves = elements.len() // batch_size
        if n_leaves & (n_leaves - 1) or n_leaves == 0:
            raise InputValidation("PowerOfTwoLengthRequired")
        n_nodes = 2 * n_leaves - 1
        nodes = hal.dev_alloc(2 * n_nodes)
        hal._check(hal._lib.b200_merkle_build(hal._ctx, elements.ptr, elements.len(), batch_size, nodes.ptr, n_nodes))
        return cls(hal, n_leaves.bit_length() - 1, nodes)

    def _digests(self, start: int, count: int) -> List[bytes]:
        raw = self.hal.to_host(self.nodes.slice(2 * start, 2 * (start + count)))
        b = np.ascontiguousarray(raw).view(np.uint8).reshape(count, 32)
        return [bytes(r) for r in b]

    def root(self) -> bytes:
        return self._digests((2 << self.log_len) - 2, 1)[0]

    def layer(self, layer_depth: int) -> List[bytes]:
        """binary_merkle_tree.rs:112-119"""
        if layer_depth > self.log_len:
            raise InputValidation("IncorrectLayerDepth")
        n_nodes = (2 << self.log_len) - 1
        start = n_nodes + 1 - (1 << (layer_depth + 1))
        return self._digests(start, 1 << layer_depth)

    def branch(self, index: int, layer_depth: int) -> List[bytes]:
        """Merkle branch of leaf `index` up to `layer_depth` (binary_merkle_tree.rs:124-141)"""
        if index >= 1 << self.log_len or layer_depth > self.log_len:
            raise InputValidation("IndexOutOfRange")
        out = []
        for j in range(self.log_len - layer_depth):
            node = (((1 << j) - 1) << (self.log_len + 1 - j)) | ((index >> j) ^ 1)
            out.append(self._digests(node, 1)[0])
        return out


def compress_pair(left: bytes, right: bytes, hal: B200Layer) -> bytes:
    """Groestl256ByteCompression on the device (one pair; verification-side helper for tests)"""
    buf = np.frombuffer(left + right, dtype=np.uint64).reshape(4, 2).copy()
    d = hal.to_device(buf)
    out = hal.dev_alloc(2)
    hal._check(hal._lib.b200_groestl256_compress_pairs(hal._ctx, d.ptr, 1, out.ptr))
    return bytes(np.ascontiguousarray(hal.to_host(out)).view(np.uint8).reshape(32))
