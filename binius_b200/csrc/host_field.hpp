// host_field.hpp -- host-side binary tower field arithmetic used by the product library for
// one-off scalar work (twiddle generation, 8-bit table generation, batch-coefficient powers).
// Independent of oracle/ (the oracle is test infrastructure and is never linked here).
//
// Field definition: reference crates/field/src/binary_field.rs:33-40, 503-527 and
// crates/field/src/arch/portable/pairwise_recursive_arithmetic.rs:12-81:
//   T_0 = GF(2),  T_{k+1} = T_k[X_k] / (X_k^2 + X_k * X_{k-1} + 1),  X_{-1} = 1,
//   element of T_k = integer of 2^k bits = lo + hi * X_{k-1}.
#pragma once
#include <cstdint>

namespace b200 {
namespace hostf {

typedef unsigned __int128 u128;

template <int K>
struct Tw {
	static constexpr int H = 1 << (K - 1);
	static inline u128 lo(u128 a) { return a & ((((u128)1) << H) - 1); }
	static inline u128 hi(u128 a) { return (a >> H) & ((((u128)1) << H) - 1); }
	// multiply by X_{K-1}
	static u128 mul_alpha(u128 a) {
		u128 a0 = lo(a), a1 = hi(a);
		return a1 | ((a0 ^ Tw<K - 1>::mul_alpha(a1)) << H);
	}
	static u128 mul(u128 a, u128 b) {
		u128 a0 = lo(a), a1 = hi(a), b0 = lo(b), b1 = hi(b);
		u128 p0 = Tw<K - 1>::mul(a0, b0);
		u128 p2 = Tw<K - 1>::mul(a1, b1);
		u128 pm = Tw<K - 1>::mul(a0 ^ a1, b0 ^ b1);
		return (p0 ^ p2) | ((pm ^ p0 ^ p2 ^ Tw<K - 1>::mul_alpha(p2)) << H);
	}
	static u128 square(u128 a) { return mul(a, a); }
	static u128 invert(u128 a) {
		u128 a0 = lo(a), a1 = hi(a);
		u128 t = a0 ^ Tw<K - 1>::mul_alpha(a1);
		u128 delta = Tw<K - 1>::mul(a0, t) ^ Tw<K - 1>::square(a1);
		u128 di = Tw<K - 1>::invert(delta);
		return Tw<K - 1>::mul(di, t) | (Tw<K - 1>::mul(di, a1) << H);
	}
};
template <>
struct Tw<0> {
	static u128 mul_alpha(u128 a) { return a & 1; }
	static u128 mul(u128 a, u128 b) { return a & b & 1; }
	static u128 square(u128 a) { return a & 1; }
	static u128 invert(u128 a) { return a & 1; }
};

inline u128 mul(u128 a, u128 b, int k) {
	switch (k) {
	case 0: return Tw<0>::mul(a, b);
	case 1: return Tw<1>::mul(a, b);
	case 2: return Tw<2>::mul(a, b);
	case 3: return Tw<3>::mul(a, b);
	case 4: return Tw<4>::mul(a, b);
	case 5: return Tw<5>::mul(a, b);
	case 6: return Tw<6>::mul(a, b);
	default: return Tw<7>::mul(a, b);
	}
}
inline u128 invert(u128 a, int k) {
	switch (k) {
	case 0: return Tw<0>::invert(a);
	case 1: return Tw<1>::invert(a);
	case 2: return Tw<2>::invert(a);
	case 3: return Tw<3>::invert(a);
	case 4: return Tw<4>::invert(a);
	case 5: return Tw<5>::invert(a);
	case 6: return Tw<6>::invert(a);
	default: return Tw<7>::invert(a);
	}
}
// Table-driven variant for hot host paths (batch-coefficient powers: one product per composition and
// round): the 8-bit level comes from a 64 KiB product table built once from the bit-level recursion,
// 81 lookups per B128 product (~0.2 us) instead of 3^7 bit products (~3.5 us).
struct Tab8 {
	uint8_t mul[65536], alpha[256];
	Tab8() {
		for (int a = 0; a < 256; a++) {
			for (int b = a; b < 256; b++) mul[(a << 8) | b] = mul[(b << 8) | a] = (uint8_t)Tw<3>::mul((u128)a, (u128)b);
			alpha[a] = (uint8_t)Tw<3>::mul_alpha((u128)a);
		}
	}
};
inline const Tab8 &tab8() {
	static const Tab8 t;
	return t;
}
template <int K>
struct FastTw {
	static constexpr int H = 1 << (K - 1);
	static inline u128 mul_alpha(u128 a) {
		u128 a0 = Tw<K>::lo(a), a1 = Tw<K>::hi(a);
		return a1 | ((a0 ^ FastTw<K - 1>::mul_alpha(a1)) << H);
	}
	static inline u128 mul(u128 a, u128 b) {
		u128 a0 = Tw<K>::lo(a), a1 = Tw<K>::hi(a), b0 = Tw<K>::lo(b), b1 = Tw<K>::hi(b);
		u128 p0 = FastTw<K - 1>::mul(a0, b0);
		u128 p2 = FastTw<K - 1>::mul(a1, b1);
		u128 pm = FastTw<K - 1>::mul(a0 ^ a1, b0 ^ b1);
		return (p0 ^ p2) | ((pm ^ p0 ^ p2 ^ FastTw<K - 1>::mul_alpha(p2)) << H);
	}
};
template <>
struct FastTw<3> {
	static inline u128 mul_alpha(u128 a) { return tab8().alpha[(uint8_t)a]; }
	static inline u128 mul(u128 a, u128 b) { return tab8().mul[((uint32_t)(uint8_t)a << 8) | (uint8_t)b]; }
};
inline u128 mul128(u128 a, u128 b) { return FastTw<7>::mul(a, b); }
// BinaryField128bPolyval product on stored (Montgomery) forms: a * b * X^-128 mod X^128 + X^127 + X^126 + X^121 + 1
// (reference crates/field/src/arch/portable/packed_polyval_128.rs:88-122).  Bit-serial carry-less product and
// bit-serial Montgomery reduction: whenever bit i of the 256-bit product is set, adding p << i clears it (p has
// constant term 1), and the result is the upper half.
inline u128 polyval_mul(u128 a, u128 b) {
	u128 lo = 0, hi = 0;
	for (int i = 0; i < 128; i++)
		if ((b >> i) & 1) {
			lo ^= a << i;
			if (i) hi ^= a >> (128 - i);
		}
	const u128 p_lo = ((u128)0xC200000000000000ull << 64) | 1ull;  // X^127 + X^126 + X^121 + 1 (X^128 is the carry into `hi`)
	for (int i = 0; i < 128; i++)
		if ((lo >> i) & 1) {
			lo ^= p_lo << i;
			hi ^= (i ? p_lo >> (128 - i) : (u128)0) ^ ((u128)1 << i);  // p << i spills into the upper half; X^128 << i = bit i of hi
		}
	return hi;
}
inline u128 from_words(const uint64_t w[2]) { return ((u128)w[1] << 64) | w[0]; }

}  // namespace hostf
}  // namespace b200
