/*
 * oracle/tower.h -- TEST INFRASTRUCTURE ONLY (not shipped, not on the product path).
 *
 * CPU restatement of the binary tower field arithmetic of IrreducibleOSS/binius.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or load anything under oracle/.
 *
 * Reference (all paths relative to /root/reference):
 *   crates/field/src/binary_field.rs:503-527, 681-698   element layout: lo | hi << 2^i, subfield = hi == 0
 *   crates/field/src/arch/portable/pairwise_recursive_arithmetic.rs:12-30  mul  (Karatsuba tower step)
 *   ... :33-45  square,  :48-62  mul_alpha,  :65-81  invert_or_zero
 *   crates/field/src/binary_field.rs:363-414   F x subfield = limb-wise product, F x B1 = mask
 *
 * T_0 = GF(2); T_{i+1} = T_i[X_i]/(X_i^2 + X_i*X_{i-1} + 1), X_{-1} := 1.
 * An element of T_k (k = log2 of the bit width) is an integer of 2^k bits: lo + hi * X_{k-1}.
 *
 * Parity pinning: see tests/test_oracle_field.py -- every multiplication KAT of
 * binary_field.rs:925-1028, the MULTIPLICATIVE_GENERATOR order test (:1031-1102) for all 8
 * fields, and the tower<->AES isomorphism constants of aes_field.rs are checked against this file.
 */
#ifndef BINIUS_ORACLE_TOWER_H
#define BINIUS_ORACLE_TOWER_H

#include <stdint.h>

typedef unsigned __int128 u128;

/* 8-bit base tables, filled by tower_init() from the bit-level recursion below. */
extern uint8_t TOWER_MUL8[256][256];
extern uint8_t TOWER_INV8[256];
extern uint8_t TOWER_ALPHA8[256]; /* x -> x * X_2 (= 0x10) in T_3 */

void tower_init(void);

/* bit-level recursive definitions (slow; the spec) */
u128 tower_mul_slow(u128 a, u128 b, int k);
u128 tower_mul_alpha_slow(u128 a, int k);

/* table-backed (fast) versions used by the op restatements; k in [0,7] */
u128 tower_mul(u128 a, u128 b, int k);
u128 tower_mul_alpha(u128 a, int k);
u128 tower_square(u128 a, int k);
u128 tower_invert(u128 a, int k); /* 0 -> 0 */

/* a in T_7 (B128), s in T_k: multiply every 2^k-bit limb of a by s  (binary_field.rs:363-393) */
u128 tower_mul_subfield(u128 a, u128 s, int k);

static inline u128 b128_mul(u128 a, u128 b) { return tower_mul(a, b, 7); }

#endif
