/*
 * oracle/groestl.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Groestl-256 written from the specification (Gauravaram et al., "Groestl -- a SHA-3 candidate", sections 3.2-3.4):
 * byte-oriented AddRoundConstant / SubBytes / ShiftBytes / MixBytes on the 8 x 8 state, S-box = AES S-box computed from
 * the GF(2^8) inverse + affine map.  This is the hash behind the reference's Merkle commitments:
 *   crates/hash/src/groestl/digest.rs:60-90        Groestl256 (IV = 0..0 || 0x0100, len64 big-endian padding, output
 *                                                  transformation P(h) ^ h truncated to the last 32 bytes)
 *   crates/hash/src/groestl/compression.rs:22-36   Groestl256ByteCompression: P(x) ^ x on the 64-byte pair of digests,
 *                                                  last 32 bytes
 *   crates/core/src/merkle_tree/binary_merkle_tree.rs:27-211   leaves = digest of each chunk of `batch_size` elements
 *                                                  (16-byte little-endian canonical serialisation), inner nodes = layers of
 *                                                  pair compressions, flattened with the root last
 * The reference pins its implementation against the groestl crate (groestl/tests.rs); this file is pinned against the
 * published known answers of Groestl-256 ("" and "abc") in tests/test_groestl_cpu.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint8_t SBOX[256];
static int sbox_ready = 0;

static uint8_t gmul(uint8_t a, uint8_t b) { /* GF(2^8) mod x^8 + x^4 + x^3 + x + 1 */
	uint8_t r = 0;
	while (b) {
		if (b & 1) r ^= a;
		a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1B : 0));
		b >>= 1;
	}
	return r;
}
static void sbox_init(void) {
	if (sbox_ready) return;
	for (int x = 0; x < 256; x++) {
		uint8_t inv = 0;
		if (x)
			for (int y = 1; y < 256; y++)
				if (gmul((uint8_t)x, (uint8_t)y) == 1) {
					inv = (uint8_t)y;
					break;
				}
		uint8_t s = inv;
		for (int k = 1; k <= 4; k++) s ^= (uint8_t)((inv << k) | (inv >> (8 - k)));
		SBOX[x] = s ^ 0x63;
	}
	sbox_ready = 1;
}

/* state[row][col]; bytes of a 64-byte block map column by column: byte 8*col + row */
static void permutation(uint8_t st[8][8], int is_q) {
	static const int SHIFT_P[8] = {0, 1, 2, 3, 4, 5, 6, 7}, SHIFT_Q[8] = {1, 3, 5, 7, 0, 2, 4, 6};
	static const uint8_t MIX[8] = {2, 2, 3, 4, 5, 3, 5, 7};
	const int *sh = is_q ? SHIFT_Q : SHIFT_P;
	for (int r = 0; r < 10; r++) {
		/* AddRoundConstant */
		if (!is_q) {
			for (int c = 0; c < 8; c++) st[0][c] ^= (uint8_t)((c << 4) ^ r);
		} else {
			for (int row = 0; row < 8; row++)
				for (int c = 0; c < 8; c++) st[row][c] ^= 0xFF;
			for (int c = 0; c < 8; c++) st[7][c] ^= (uint8_t)((c << 4) ^ r);
		}
		/* SubBytes + ShiftBytes (row i rotated left by sh[i]) */
		uint8_t t[8][8];
		for (int row = 0; row < 8; row++)
			for (int c = 0; c < 8; c++) t[row][c] = SBOX[st[row][(c + sh[row]) & 7]];
		/* MixBytes: column <- circ(02,02,03,04,05,03,05,07) * column */
		for (int c = 0; c < 8; c++)
			for (int row = 0; row < 8; row++) {
				uint8_t v = 0;
				for (int k = 0; k < 8; k++) v ^= gmul(MIX[k], t[(row + k) & 7][c]);
				st[row][c] = v;
			}
	}
}
static void load(uint8_t st[8][8], const uint8_t *b) {
	for (int c = 0; c < 8; c++)
		for (int r = 0; r < 8; r++) st[r][c] = b[8 * c + r];
}
static void store(const uint8_t st[8][8], uint8_t *b) {
	for (int c = 0; c < 8; c++)
		for (int r = 0; r < 8; r++) b[8 * c + r] = st[r][c];
}
static void compress(uint8_t h[64], const uint8_t m[64]) {
	uint8_t p[8][8], q[8][8], hm[64], pb[64], qb[64];
	for (int i = 0; i < 64; i++) hm[i] = h[i] ^ m[i];
	load(p, hm);
	load(q, m);
	permutation(p, 0);
	permutation(q, 1);
	store(p, pb);
	store(q, qb);
	for (int i = 0; i < 64; i++) h[i] ^= pb[i] ^ qb[i];
}

void orc_groestl256(const uint8_t *msg, uint64_t len, uint8_t out[32]) {
	sbox_init();
	uint8_t h[64] = {0};
	h[62] = 0x01; /* IV: output length 256 as a big-endian 64-bit integer in the last column */
	uint64_t off = 0;
	for (; off + 64 <= len; off += 64) compress(h, msg + off);
	uint8_t last[128] = {0};
	uint64_t rem = len - off;
	memcpy(last, msg + off, rem);
	last[rem] = 0x80;
	uint64_t n_blocks = len / 64 + (rem <= 55 ? 1 : 2);
	uint64_t tail = rem <= 55 ? 64 : 128;
	for (int i = 0; i < 8; i++) last[tail - 1 - i] = (uint8_t)(n_blocks >> (8 * i));
	compress(h, last);
	if (tail == 128) compress(h, last + 64);
	uint8_t p[8][8], pb[64];
	load(p, h);
	permutation(p, 0);
	store(p, pb);
	for (int i = 0; i < 32; i++) out[i] = pb[32 + i] ^ h[32 + i];
}

/* Groestl256ByteCompression: last 32 bytes of P(x) ^ x, x = left || right */
void orc_groestl256_compress_pair(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]) {
	sbox_init();
	uint8_t x[64], pb[64], p[8][8];
	memcpy(x, left, 32);
	memcpy(x + 32, right, 32);
	load(p, x);
	permutation(p, 0);
	store(p, pb);
	for (int i = 0; i < 32; i++) out[i] = pb[32 + i] ^ x[32 + i];
}

/* BinaryMerkleTree::build over `n_leaves` chunks of `leaf_bytes` bytes: nodes = leaves layer, then each layer of pair
 * compressions, root last: (2 * n_leaves - 1) digests of 32 bytes */
void orc_merkle_build(const uint8_t *data, uint64_t n_leaves, uint64_t leaf_bytes, uint8_t *nodes) {
	for (uint64_t i = 0; i < n_leaves; i++) orc_groestl256(data + i * leaf_bytes, leaf_bytes, nodes + 32 * i);
	uint8_t *prev = nodes, *cur = nodes + 32 * n_leaves;
	for (uint64_t n = n_leaves / 2; n >= 1; n /= 2) {
		for (uint64_t i = 0; i < n; i++) orc_groestl256_compress_pair(prev + 64 * i, prev + 64 * i + 32, cur + 32 * i);
		prev = cur;
		cur += 32 * n;
	}
}
