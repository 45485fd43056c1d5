/*
 * binius_b200.h -- C ABI of the B200-native (sm_100a) implementation of the Binius prover hot path.
 *
 * This is the drop-in boundary: the entry points below are what a thin Rust `cc`/bindgen shim
 * implementing the reference's own traits would bind:
 *
 *   binius_compute::ComputeLayer / ComputeLayerExecutor / KernelExecutor
 *        (reference crates/compute/src/layer.rs:22-88, 100-510, 518-590)
 *   binius_ntt::AdditiveNTT                       (crates/ntt/src/additive_ntt.rs:58-166)
 *   binius_hal::ComputationBackend                (crates/hal/src/backend.rs:35-83)
 *
 * Conventions
 *   - A field element F = BinaryField128b is a little-endian u128, passed as 2 x uint64 {lo, hi}
 *     (crates/field/src/binary_field.rs:115-133).  Element i of a slice lives at byte offset 16*i.
 *   - `b200_dev_ptr` values are DEVICE pointers (element-aligned, 16 B); an FSlice/FSliceMut of the
 *     reference maps to (b200_dev_ptr, n_elems).  ComputeMemory::ALIGNMENT is 1: any element
 *     offset is a valid slice start (crates/compute/src/memory.rs:69-235).
 *   - Scalars, challenge vectors and index arrays are HOST pointers.
 *   - Every call returns a status; B200_OK = 0.  Statuses map onto compute::Error
 *     (crates/compute/src/layer.rs:706-716) and ntt::Error (crates/ntt/src/error.rs:3-29).
 *   - All work of one context is issued on one CUDA stream in call order, so the store-to-load
 *     ordering contract of ComputeLayerExecutor (layer.rs:90-99) holds.  Calls are asynchronous
 *     unless stated; scalar results are deferred "OpValue" slots read back by b200_results_fetch.
 *     Two kinds of call are DEFERRED inside the library (layer.rs:90-99 allows it): consecutive
 *     b200_extrapolate_line calls with one challenge are queued and launched together by the next call of
 *     any other entry point (b200_ctx_stream, b200_event_record and b200_flush included), and the ops
 *     of an open kernel scope are lowered at b200_kernel_scope_end.
 *   - A context may be shared by several host threads (the trait methods take `&self` and `join`/`map`
 *     closures run on rayon threads, layer.rs:115-131): every entry point takes the context's lock and
 *     makes its device current for the call; a kernel scope holds the lock from begin to end.  The result
 *     slots are per context and b200_results_reset starts a new numbering, so an `execute` (reset .. fetch) is a
 *     scope the HOST side must not interleave between threads: the shim holds a mutex per layer around it, as
 *     binius_b200/layer.py and host/compute_layer.hpp do (tests/test_gpu_layer.py::
 *     test_one_layer_from_several_host_threads).
 *   - There is no CPU fallback: every entry point fails with B200_ERR_DEVICE when no sm_100 device
 *     is usable.
 */
#ifndef BINIUS_B200_H
#define BINIUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;
typedef void *b200_dev_ptr;

enum {
	B200_OK = 0,
	B200_ERR_INPUT_VALIDATION = 1, /* compute::Error::InputValidation */
	B200_ERR_ALLOC = 2,            /* compute::Error::Alloc(OutOfMemory) */
	B200_ERR_DEVICE = 3,           /* compute::Error::DeviceError */
	/* ntt::Error classes (crates/ntt/src/error.rs) */
	B200_ERR_NTT_POWER_OF_TWO = 11,  /* PowerOfTwoLengthRequired */
	B200_ERR_NTT_SKIP_ROUNDS = 12,   /* SkipRoundsTooLarge */
	B200_ERR_NTT_BATCH = 13,         /* BatchTooLarge */
	B200_ERR_NTT_COSET = 14,         /* CosetIndexOutOfBounds */
	B200_ERR_NTT_DOMAIN = 15,        /* DomainTooSmall */
	B200_ERR_NTT_FIELD = 16          /* FieldTooSmall */
};

/* ---- context / memory (ComputeHolder + ComputeLayer::copy_*, layer.rs:36-55, 732-776) ---------- */
int32_t b200_ctx_create(int32_t device, b200_ctx **out);
void b200_ctx_destroy(b200_ctx *ctx);
const char *b200_last_error(b200_ctx *ctx);
/* the CUDA stream of the context (cudaStream_t as void*) so callers can record events on it; launches
 * every queued fold first, so work the caller enqueues on the stream afterwards is ordered behind it */
void *b200_ctx_stream(b200_ctx *ctx);
/* launch everything the library has deferred (queued folds); asynchronous */
int32_t b200_flush(b200_ctx *ctx);
/* adopt an external stream (e.g. torch.cuda.current_stream().cuda_stream); NULL restores the own one */
int32_t b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream);
/* Kernel-selection switches for A/B measurements and tests (never needed for correctness; the defaults are
 * the production paths).  Keys: "ntt" (0 look-up-table passes + bit-sliced low layers, 1 bit-sliced only,
 * 2 scalar tables), "ntt_log_cc" (5..7), "fold" (2 TMA-staged, 1 K64, 0 LUT128), "round_evals_tc" (1/0/2),
 * "uni_generic" (1 forces the generic univariate-skip kernel).  Unknown key: InputValidation. */
int32_t b200_ctx_set_tuning(b200_ctx *ctx, const char *key, int32_t value);
/* Host-side scalar product in BinaryField128b (no device involved): the O(1)-per-round scalar work of a host
 * mirror without its own field arithmetic -- batch-coefficient powers (bivariate_product.rs:337-339), eq factors. */
void b200_host_mul128(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);
/* CUDA-event timing on the context's stream (bench.py): record two events, read elapsed ms */
int32_t b200_event_create(b200_ctx *ctx, void **event_out);
int32_t b200_event_record(b200_ctx *ctx, void *event);
int32_t b200_event_elapsed_ms(b200_ctx *ctx, void *start, void *stop, float *ms_out); /* syncs on stop */
void b200_event_destroy(void *event);
/* number of kernels this library has launched on the context so far (bench `gpu_launches`) */
uint64_t b200_ctx_launch_count(b200_ctx *ctx);

/* device arena handed to the reference's BumpAllocator (alloc.rs:31-105): n_elems B128 elements */
int32_t b200_dev_alloc(b200_ctx *ctx, uint64_t n_elems, b200_dev_ptr *out);
int32_t b200_dev_free(b200_ctx *ctx, b200_dev_ptr p);
/* pinned host staging memory (optional; pageable host pointers are accepted everywhere) */
int32_t b200_host_alloc(b200_ctx *ctx, uint64_t n_bytes, void **out);
int32_t b200_host_free(b200_ctx *ctx, void *p);

int32_t b200_copy_h2d(b200_ctx *ctx, const void *host_src, b200_dev_ptr dst, uint64_t n_elems);
int32_t b200_copy_d2h(b200_ctx *ctx, b200_dev_ptr src, void *host_dst, uint64_t n_elems); /* synchronous */
/* copy_h2d on the context's SIDE stream (pinned source): overlaps with the work of the main stream until
 * b200_side_join makes the main stream wait for every side copy issued so far.  For streaming a witness in while
 * earlier chunks are processed; the destination must not be used by main-stream work before the join. */
int32_t b200_copy_h2d_side(b200_ctx *ctx, const void *host_src, b200_dev_ptr dst, uint64_t n_elems);
int32_t b200_side_join(b200_ctx *ctx);
int32_t b200_copy_d2d(b200_ctx *ctx, b200_dev_ptr src, b200_dev_ptr dst, uint64_t n_elems);
/* ComputeLayer::fill (layer.rs:74-87) */
int32_t b200_fill(b200_ctx *ctx, b200_dev_ptr dst, uint64_t n_elems, const uint64_t value[2]);
int32_t b200_sync(b200_ctx *ctx);

/* ---- deferred scalar results (ComputeLayerExecutor::OpValue; ComputeLayer::execute returns them,
 *      layer.rs:62-72).  Slots are valid until b200_results_reset. ------------------------------ */
int32_t b200_results_reset(b200_ctx *ctx);
int32_t b200_results_fetch(b200_ctx *ctx, const uint32_t *slots, uint32_t n, uint64_t *host_out /* 2*n */);

/* ---- ComputeLayerExecutor ops ------------------------------------------------------------------ */
/* extrapolate_line (layer.rs:402-426; cpu/layer.rs:393-408): e0[i] += (e1[i]-e0[i])*z.
 * Deferred: the call only queues the fold; folds with the same z on disjoint slices go out as ONE
 * multi-segment launch when any other entry point (or a conflicting fold) arrives -- see b200_flush. */
int32_t b200_extrapolate_line(b200_ctx *ctx, b200_dev_ptr evals_0, uint64_t n0, b200_dev_ptr evals_1,
							  uint64_t n1, const uint64_t z[2]);
/* The same fold on HOST buffers (a ComputationBackend whose Vec<P> is host-dereferenceable,
 * hal/src/backend.rs:19-31, 65-75): chunked H2D -> kernel -> D2H pipeline over both PCIe directions.
 * Synchronous; host_e0 is updated in place.  Pinned buffers (b200_host_alloc) give full overlap. */
int32_t b200_extrapolate_line_host(b200_ctx *ctx, void *host_e0, const void *host_e1, uint64_t n,
								   const uint64_t z[2]);
/* tensor_expand (layer.rs:269-296; cpu/layer.rs:282-302) */
int32_t b200_tensor_expand(b200_ctx *ctx, b200_dev_ptr data, uint64_t data_len, uint32_t log_n,
						   const uint64_t *coordinates /* 2*k */, uint32_t k);
/* inner_product (layer.rs:247-267; cpu/layer.rs:205-236): a = SubfieldSlice{a, tower_level} */
int32_t b200_inner_product(b200_ctx *ctx, b200_dev_ptr a, uint64_t n_a, uint32_t tower_level,
						   b200_dev_ptr b, uint64_t n_b, uint32_t *result_slot);
/* fold_left / fold_right (layer.rs:298-356; cpu/layer.rs:238-280, 574-675).
 * fold_right with vec.len() * 2^tower_level == 128 (one output per matrix element: the projection of a packed column
 * by 128 / 2^level coefficients, prove/zerocheck.rs:416-434) is deferred like extrapolate_line: consecutive calls with the
 * same `vec` on disjoint buffers go out as ONE launch when any other entry point (or a conflicting fold) arrives. */
int32_t b200_fold_left(b200_ctx *ctx, b200_dev_ptr mat, uint64_t n_mat, uint32_t tower_level,
					   b200_dev_ptr vec, uint64_t n_vec, b200_dev_ptr out, uint64_t n_out);
int32_t b200_fold_right(b200_ctx *ctx, b200_dev_ptr mat, uint64_t n_mat, uint32_t tower_level,
						b200_dev_ptr vec, uint64_t n_vec, b200_dev_ptr out, uint64_t n_out);

/* ---- Groestl-256 Merkle commitments over device-resident data (commit pipeline) ----------
 * A digest is 32 bytes = 2 B128 slots of device memory.  Leaves: digests[i] = Groestl256 of leaf i's `leaf_elems`
 * B128 elements in their 16-byte little-endian serialisation (hash_interleaved, crates/core/src/merkle_tree/
 * binary_merkle_tree.rs:170-211 with crates/hash/src/groestl/digest.rs:60-90); pairs: out[i] =
 * Groestl256ByteCompression(in[2i], in[2i+1]) (crates/hash/src/groestl/compression.rs:22-36);
 * b200_merkle_build = BinaryMerkleTree::build (binary_merkle_tree.rs:27-102): `nodes` receives the leaf layer followed
 * by every inner layer, root last (n_nodes = 2 * n_elems / batch_size - 1 digests).  Errors mirror
 * merkle_tree/errors.rs: IncorrectBatchSize, PowerOfTwoLengthRequired, IncorrectVectorLen. */
int32_t b200_groestl256_leaves(b200_ctx *ctx, b200_dev_ptr data, uint64_t n_leaves, uint64_t leaf_elems, b200_dev_ptr digests);
int32_t b200_groestl256_compress_pairs(b200_ctx *ctx, b200_dev_ptr in_digests, uint64_t n_pairs, b200_dev_ptr out_digests);
int32_t b200_merkle_build(b200_ctx *ctx, b200_dev_ptr elements, uint64_t n_elems, uint64_t batch_size, b200_dev_ptr nodes, uint64_t n_nodes);

/* ---- GF(2)-linear maps on B128 and the POLYVAL ("fast") field ---------------------------------------------
 * dst[i] = L(src[i]), L given by its 128 basis images (2 words each): FieldLinearTransformation::transform
 * (crates/field/src/linear_transformation.rs), the device side of convert_witnesses_to_fast_ext
 * (core/src/constraint_system/prove.rs:291-292) with the tables of crates/field/src/polyval.rs:516-788.
 * In place (dst == src) allowed.  The GKR grand-product prover (core/src/protocols/gkr_gpa) runs in
 * BinaryField128bPolyval on the CPU because CLMUL makes that field cheap there; it is ISOMORPHIC to the tower
 * field, so on this device the layers and sumcheck rounds run on the tower kernels and only what crosses the
 * interface is mapped: phi(a * b) = phi(a) * phi(b) bit for bit (tests pin this against the Montgomery arithmetic). */
int32_t b200_linear_map(b200_ctx *ctx, b200_dev_ptr src, b200_dev_ptr dst, uint64_t n_elems, const uint64_t *basis_images /* 2*128 */);
/* host-side: product of two BinaryField128bPolyval elements in stored (Montgomery) form
 * (crates/field/src/arch/portable/packed_polyval_128.rs:88-122) */
void b200_host_polyval_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);
/* host-side: the tower <-> POLYVAL basis change (128 images each way), derived from the two published generators */
int32_t b200_host_polyval_basis_change(uint64_t tower_to_polyval[256], uint64_t polyval_to_tower[256]);

/* Arithmetic circuit (crates/math/src/arith_expr.rs:200-206); ComputeLayer::compile_expr */
typedef struct {
	uint32_t op; /* 0 Add(l,r)  1 Mul(l,r)  2 Pow(l, r = exponent)  3 Const(c)  4 Var(l) */
	uint32_t l;
	uint64_t r;
	uint64_t c_lo, c_hi;
} b200_expr_step;
typedef struct b200_expr b200_expr;
int32_t b200_expr_compile(b200_ctx *ctx, const b200_expr_step *steps, uint32_t n_steps, b200_expr **out);
void b200_expr_free(b200_expr *e);
uint32_t b200_expr_n_vars(const b200_expr *e);

/* compute_composite (layer.rs:428-464; cpu/layer.rs:410-435) */
int32_t b200_compute_composite(b200_ctx *ctx, const b200_dev_ptr *inputs, uint32_t n_inputs,
							   uint64_t row_len, b200_dev_ptr out, uint64_t n_out, const b200_expr *expr);
/* pairwise_product_reduce (layer.rs:466-509; cpu/layer.rs:437-484) */
int32_t b200_pairwise_product_reduce(b200_ctx *ctx, b200_dev_ptr input, uint64_t n_in,
									 const b200_dev_ptr *round_outputs, const uint64_t *round_output_lens,
									 uint32_t n_rounds);

/* ---- accumulate_kernels / map_kernels (layer.rs:134-245) and KernelExecutor ops (layer.rs:518-590) ------
 * The layer's accumulate_kernels(map, mem_maps) = b200_kernel_scope_begin; one b200_kernel_local per
 * KernelMemMap::Local (layer.rs:617-644; zero-initialised scratch that lives until the scope ends); the
 * caller's closure with log_chunks = 0 (whole buffers) issuing the four ops below; b200_kernel_scope_end.
 * Inside a scope the ops are recorded and lowered together ("free to call the specification closure",
 * layer.rs:176-177; a Value "is a promise", layer.rs:522): sums of degree <= 2 expressions over inputs that
 * are mapped buffers or Locals written by one add(src1, src2) become tensor-core inner-product jobs on the
 * (src1, src2) pointer pairs and the Locals are never materialised -- the round evaluation of
 * core/src/protocols/sumcheck/v3/bivariate_product.rs:343-405; anything else runs op by op as recorded.
 * Outside a scope each op launches at once. */
int32_t b200_kernel_scope_begin(b200_ctx *ctx);
int32_t b200_kernel_local(b200_ctx *ctx, uint32_t log_size, b200_dev_ptr *out);
int32_t b200_kernel_scope_end(b200_ctx *ctx);
/* decl_value */
int32_t b200_kernel_decl_value(b200_ctx *ctx, const uint64_t init[2], uint32_t *value_slot);
/* sum_composition_evals: slot += batch_coeff * sum_i expr(inputs[.][i]) */
int32_t b200_kernel_sum_composition_evals(b200_ctx *ctx, const b200_dev_ptr *inputs, uint32_t n_inputs,
										  uint64_t row_len, const b200_expr *expr,
										  const uint64_t batch_coeff[2], uint32_t value_slot);
int32_t b200_kernel_add(b200_ctx *ctx, uint32_t log_len, b200_dev_ptr src1, b200_dev_ptr src2, b200_dev_ptr dst);
int32_t b200_kernel_add_assign(b200_ctx *ctx, uint32_t log_len, b200_dev_ptr src, b200_dev_ptr dst);

/* Fused form of v3::calculate_round_evals (core/src/protocols/sumcheck/v3/bivariate_product.rs:
 * 303-408): the traced accumulate_kernels program {sum(hi_a*hi_b), add(lo,hi), sum(inf_a*inf_b)}
 * over m multilinears of 2^n_vars elements; writes y_1 and y_inf to two result slots. */
int32_t b200_bivariate_round_evals(b200_ctx *ctx, const b200_dev_ptr *multilins, uint32_t n_multilins,
								   uint32_t n_vars, const uint32_t *idx_a, const uint32_t *idx_b,
								   uint32_t n_compositions, const uint64_t batch_coeff[2],
								   uint32_t *slot_y1, uint32_t *slot_yinf);

/* ---- additive NTT (binius_ntt::AdditiveNTT<F>, additive_ntt.rs:58-166) ------------------------- */
typedef struct b200_ntt b200_ntt;
/* SingleThreadedNTT::new(log_domain_size) over BinaryField{8,16,32}b: field_log_bits = 3,4,5
 * (single_threaded.rs:27-45; twiddles per twiddle.rs:244-313) */
int32_t b200_ntt_create(b200_ctx *ctx, uint32_t field_log_bits, uint32_t log_domain_size, b200_ntt **out);
void b200_ntt_destroy(b200_ntt *ntt);
uint32_t b200_ntt_log_domain_size(const b200_ntt *ntt);
/* get_subspace_eval(i, j) (single_threaded.rs:91-93), host-side */
int32_t b200_ntt_get_subspace_eval(const b200_ntt *ntt, uint32_t i, uint64_t j, uint64_t out[2]);
/* forward/inverse_transform on DEVICE data of n_elems scalars of 2^elem_log_bits bits each
 * (elem_log_bits >= field_log_bits; 7 = packed B128 i.e. *_transform_ext, additive_ntt.rs:137-165).
 * Scalar index = x | y << log_x | z << (log_x+log_y)  (additive_ntt.rs:8-27). */
int32_t b200_ntt_forward(b200_ctx *ctx, const b200_ntt *ntt, b200_dev_ptr data, uint32_t elem_log_bits,
						 uint64_t n_elems, uint32_t log_x, uint32_t log_y, uint32_t log_z, uint64_t coset,
						 uint32_t coset_bits, uint32_t skip_rounds);
int32_t b200_ntt_inverse(b200_ctx *ctx, const b200_ntt *ntt, b200_dev_ptr data, uint32_t elem_log_bits,
						 uint64_t n_elems, uint32_t log_x, uint32_t log_y, uint32_t log_z, uint64_t coset,
						 uint32_t coset_bits, uint32_t skip_rounds);
/* The trait-shaped variants: data is a HOST `&mut [P]`; staged H2D, transformed, D2H (synchronous). */
int32_t b200_ntt_forward_host(b200_ctx *ctx, const b200_ntt *ntt, void *host_data, uint32_t elem_log_bits,
							  uint64_t n_elems, uint32_t log_x, uint32_t log_y, uint32_t log_z,
							  uint64_t coset, uint32_t coset_bits, uint32_t skip_rounds);
int32_t b200_ntt_inverse_host(b200_ctx *ctx, const b200_ntt *ntt, void *host_data, uint32_t elem_log_bits,
							  uint64_t n_elems, uint32_t log_x, uint32_t log_y, uint32_t log_z,
							  uint64_t coset, uint32_t coset_bits, uint32_t skip_rounds);

/* fri_fold (layer.rs:358-400; cpu/layer.rs:304-391) */
int32_t b200_fri_fold(b200_ctx *ctx, const b200_ntt *ntt, uint32_t log_len, uint32_t log_batch_size,
					  const uint64_t *challenges /* 2*n_challenges */, uint32_t n_challenges,
					  b200_dev_ptr data_in, uint64_t n_in, b200_dev_ptr data_out, uint64_t n_out);

/* ---- ComputationBackend (old HAL, crates/hal/src/backend.rs:35-83) on device-resident data ---- */
/* tensor_product_full_query(query) (backend.rs:42-45 -> math/src/tensor_prod_eq_ind.rs:94-101):
 * out[0..2^k) = eq-indicator expansion of `query`; out must hold 2^k elements. */
int32_t b200_tensor_product_full_query(b200_ctx *ctx, const uint64_t *query /* 2*k */, uint32_t k,
									   b200_dev_ptr out, uint64_t n_out);
/* sumcheck_fold_multilinears, HighToLow, Folded branch (backend.rs:65-75 ->
 * hal/src/sumcheck_folding.rs:149-241 -> math/src/fold.rs:648-696 fold_left_lerp_inplace):
 * each multilinear of 2^n_vars evals with `non_const_prefix[t]` stored elements followed by an
 * implicit constant suffix `suffix_eval[t]`; folds in place and reports the new stored length. */
int32_t b200_fold_multilinears_high_to_low(b200_ctx *ctx, const b200_dev_ptr *multilins,
										   uint32_t n_multilins, uint32_t n_vars,
										   const uint64_t *non_const_prefix, const uint64_t *suffix_evals /* 2*m */,
										   const uint64_t challenge[2], uint64_t *new_lens);
/* sumcheck_compute_round_evals for eq-ind (zerocheck) provers, HighToLow (backend.rs:48-62 ->
 * hal/src/sumcheck_round_calculation.rs:126-349; evaluator core/.../prove/eq_ind.rs:646-731):
 * for each composition c and each requested evaluation point, R = sum_i E[i] * C_z(P(i)).
 * eval point codes: 1 = evaluate at 1, 2 = Karatsuba infinity point (leading term on hi-lo),
 * k >= 3 = finite domain point given in `domain_points`.  Results: n_compositions * n_points slots.
 * Folded multilinears may be truncated: element i >= stored_lens[t] reads as suffix_evals[t]
 * (sumcheck_round_calculation.rs:573-600); the reference's const-suffix shortcut (eq_ind.rs:704-721)
 * is an analytic form of the same sum. */
int32_t b200_eq_ind_round_evals(b200_ctx *ctx, const b200_dev_ptr *multilins,
								const uint64_t *stored_lens /* NULL = all 2^n_vars */,
								const uint64_t *suffix_evals /* 2*m, NULL = zeros */, uint32_t n_multilins,
								uint32_t n_vars, b200_dev_ptr eq_ind /* 2^(n_vars-1) */,
								const b200_expr *const *compositions,
								const b200_expr *const *compositions_leading, uint32_t n_compositions,
								const uint32_t *point_codes, const uint64_t *domain_points /* 2*n_points */,
								uint32_t n_points, uint32_t *first_slot);

/* EvaluationOrder (crates/math/src/multilinear_query.rs / hal/src/sumcheck_folding.rs:16-35) */
enum { B200_LOW_TO_HIGH = 0, B200_HIGH_TO_LOW = 1 };

/* sumcheck_compute_round_evals in either evaluation order and for both evaluator kinds
 * (hal/src/sumcheck_round_calculation.rs:126-349; LowToHighAccess :408-504 pairs (2i, 2i+1),
 * HighToLowAccess :507-604 pairs (i, i + 2^(n-1))).  eq_ind != NULL: the eq-ind evaluator
 * (core/.../prove/eq_ind.rs:646-731, sums weighted by eq_ind[i]); eq_ind == NULL: the regular evaluator
 * (core/.../prove/regular_sumcheck.rs:233-277, plain sums).  Other arguments as b200_eq_ind_round_evals. */
int32_t b200_sumcheck_round_evals(b200_ctx *ctx, uint32_t evaluation_order, const b200_dev_ptr *multilins,
								  const uint64_t *stored_lens, const uint64_t *suffix_evals, uint32_t n_multilins,
								  uint32_t n_vars, b200_dev_ptr eq_ind /* 2^(n_vars-1) or NULL */,
								  const b200_expr *const *compositions,
								  const b200_expr *const *compositions_leading, uint32_t n_compositions,
								  const uint32_t *point_codes, const uint64_t *domain_points, uint32_t n_points,
								  uint32_t *first_slot);
/* sumcheck_fold_multilinears, LowToHigh, Folded branch (hal/src/sumcheck_folding.rs:37-147 ->
 * math/src/fold.rs:528-575 fold_right_lerp): out[i] = in[2i] + (in[2i+1] - in[2i]) * challenge over the
 * stored prefix (an odd tail pairs with suffix_eval); out-of-place like the reference (outputs must
 * not overlap inputs); new_lens[t] = ceil(prefix[t] / 2). */
int32_t b200_fold_multilinears_low_to_high(b200_ctx *ctx, const b200_dev_ptr *multilins, const b200_dev_ptr *outputs,
										   uint32_t n_multilins, uint32_t n_vars, const uint64_t *non_const_prefix,
										   const uint64_t *suffix_evals /* 2*m */, const uint64_t challenge[2],
										   uint64_t *new_lens);

/* ---- persistent tail of an eq-ind sumcheck ------------------------------------------------------------------
 * The last rounds of a sumcheck are a few kilobytes of data: through separate calls every round pays launches, an
 * argument copy and a synchronising read (~40 us).  b200_sumcheck_tail_start launches ONE kernel that runs all
 * remaining `n_vars` rounds of an eq-ind (zerocheck) prover on full-length folded multilinears (HighToLow): per round
 * it posts the round values ([composition][point], semantics of b200_sumcheck_round_evals) to a host-mapped mailbox,
 * waits for the challenge in a host-mapped mailbox, folds every multilinear in place (fold_left_lerp_inplace) and
 * halves the eq-indicator (fold_partial_eq_ind, core/src/protocols/sumcheck/prove/common.rs:60-68).  Host side per
 * round: b200_sumcheck_tail_round_evals (blocks until the values are there), then b200_sumcheck_tail_challenge;
 * after the last challenge b200_sumcheck_tail_finish.  A backend maps the trait calls of those rounds
 * (sumcheck_compute_round_evals / sumcheck_fold_multilinears, hal/src/backend.rs:48-75) onto this triple.  A watchdog
 * ends the kernel if no challenge arrives within 5 s (the calls then fail with B200_ERR_DEVICE).
 * Instances of more than 32 hypercube points (n_vars in [7, 28], at most 1024 values per round) run on a co-resident
 * grid of up to 128 CTAs (cooperative launch, csrc/tail_grid.cuh), so a WHOLE cache-sized sumcheck (BASELINE config #3)
 * is one kernel; smaller ones on a single CTA.  `first_round_skip` = number of leading entries of `point_codes` the
 * first round of the tail does not need (a zerocheck prover's first round evaluates at infinity only, its later rounds
 * at 1 and infinity: prove/eq_ind.rs:646-731); their values are posted as zero.  Grid kernel only. */
typedef struct b200_tail b200_tail;
int32_t b200_sumcheck_tail_start(b200_ctx *ctx, const b200_dev_ptr *multilins, uint32_t n_multilins, uint32_t n_vars, b200_dev_ptr eq_ind,
								 const b200_expr *const *compositions, const b200_expr *const *compositions_leading, uint32_t n_compositions,
								 const uint32_t *point_codes, const uint64_t *domain_points, uint32_t n_points, uint32_t first_round_skip,
								 b200_tail **out);
int32_t b200_sumcheck_tail_round_evals(b200_tail *tail, uint64_t *host_out /* 2 * n_compositions * n_points */);
int32_t b200_sumcheck_tail_challenge(b200_tail *tail, const uint64_t challenge[2]);
int32_t b200_sumcheck_tail_finish(b200_tail *tail);

/* ---- zerocheck univariate-skip round ----------------------------------------------------------------
 * zerocheck_univariate_evals (core/src/protocols/sumcheck/prove/univariate.rs:235-500, with ntt_extrapolate
 * :642-678, spread_product :503-563, extrapolate_round_evals :565-640), FDomain = BinaryField8b:
 *   R[c][i] = sum_{s < 2^(n_vars-skip)} eq_ind[s] * C_c(P_0(s, x_i), ..., P_{m-1}(s, x_i)),  x_i = B8(2^skip + i),
 * P_j(s, .) = the polynomial of degree < 2^skip through the sub-cube M_j[s*2^skip ..] at the B8 points
 * 0..2^skip-1.  Composition c is evaluated at its (degree_c - 1) << skip points and extended to
 * max_domain_size with zeros assumed on the skipped domain, exactly as the reference does.
 * multilins[j]: DEVICE packed sub-field multilinear of 2^n_vars scalars at tower level tower_levels[j]
 * (0, 3..7; MLEEmbeddingAdapter layout, 2^(7-level) scalars per B128 word, low limb first); the base field
 * FBase is the smallest level >= 3 holding every column and composition constant.  eq_ind: DEVICE,
 * tensor_product_full_query(zerocheck_challenges) -- the reference makes that backend call itself
 * (univariate.rs:313-314) and returns the table as partial_eq_ind_evals.
 * host_round_evals: 2 * n_compositions * (max_domain_size - 2^skip) words, [composition][point]; synchronous.
 * Errors (InputValidation): TooManySkippedRounds, IncorrectZerocheckChallengesLength (n_eq),
 * LagrangeDomainTooSmall (max_domain_size < max degree << skip), DomainSizeTooLarge (max_domain_size > 256);
 * at most 256 multilinears and 128 compositions per call. */
int32_t b200_zerocheck_univariate_evals(b200_ctx *ctx, const b200_dev_ptr *multilins, const uint32_t *tower_levels,
										uint32_t n_multilins, uint32_t n_vars, uint32_t skip_rounds,
										b200_dev_ptr eq_ind, uint64_t n_eq, const b200_expr *const *compositions,
										const uint32_t *composition_degrees, uint32_t n_compositions,
										uint32_t max_domain_size, uint64_t *host_round_evals);

/* The same round with the witness still in HOST memory (ComputeLayer::copy_h2d of the columns fused into the round):
 * host_columns[j] (pinned; same packing as the device column) is uploaded into multilins[j] in 2^log_chunks row chunks
 * on the context's side stream while the previous chunk is evaluated -- the round values are XOR-sums over sub-cubes and
 * the domain extension is linear, so the per-chunk tables add up; values are identical to upload + the call above.
 * log_chunks is reduced until a chunk of every column is whole B128 words.  Afterwards the columns are resident in
 * multilins[] for the multilinear rounds.  Same errors as above. */
int32_t b200_zerocheck_univariate_evals_streamed(b200_ctx *ctx, const void *const *host_columns, const b200_dev_ptr *multilins,
												 const uint32_t *tower_levels, uint32_t n_multilins, uint32_t n_vars, uint32_t skip_rounds,
												 b200_dev_ptr eq_ind, uint64_t n_eq, const b200_expr *const *compositions,
												 const uint32_t *composition_degrees, uint32_t n_compositions,
												 uint32_t max_domain_size, uint32_t log_chunks, uint64_t *host_round_evals);

/* The round in two halves, so that the part which needs NO verifier challenge overlaps with the witness upload and the
 * commitment (in the reference the zerocheck challenges are sampled after the commitment has been observed,
 * constraint_system/prove.rs: commit -> zerocheck; the sub-cube extrapolations and the composition values on the
 * extrapolation domain, univariate.rs:380-470, depend on the witness only).
 *   prepare: extrapolates the columns and evaluates the non-linear part of every composition at its (deg - 1) * 2^skip
 *            points for every sub-cube, as B8 values, into `store` (b200_zerocheck_univariate_store_elems B128 elements of
 *            caller-owned device memory: 2^(n_vars - skip) * n_compositions * (max degree - 1) * 2^skip bytes -- 8x the
 *            quadratically used witness bits at skip 7).  host_columns non-NULL: the columns are uploaded in
 *            2^log_chunks row chunks on the side stream as in the streamed call above and chunk c is prepared while
 *            chunk c + 1 is in flight; NULL: the columns are resident.  *prepared = 1 when the shape is covered
 *            (B1/B8 columns, B8 constants, monomials of degree <= 2 with byte coefficients, skip_rounds >= 2); 0: nothing
 *            was stored (the columns were still uploaded) and the caller runs b200_zerocheck_univariate_evals instead.
 *   finish : out[c][i] = sum_s eq[s] * store[s][c][i] + the linear monomials (evaluate_partial_high of the column by eq on
 *            the tensor cores) + the reference's domain extension; same arguments as b200_zerocheck_univariate_evals
 *            plus the store; values identical to that call.  InputValidation when the shape is not a prepared one. */
uint64_t b200_zerocheck_univariate_store_elems(uint32_t n_vars, uint32_t skip_rounds, const uint32_t *composition_degrees, uint32_t n_compositions);
int32_t b200_zerocheck_univariate_prepare(b200_ctx *ctx, const void *const *host_columns, const b200_dev_ptr *multilins,
										  const uint32_t *tower_levels, uint32_t n_multilins, uint32_t n_vars, uint32_t skip_rounds,
										  const b200_expr *const *compositions, const uint32_t *composition_degrees,
										  uint32_t n_compositions, uint32_t max_domain_size, uint32_t log_chunks,
										  b200_dev_ptr store, uint64_t store_elems, uint32_t *prepared);
int32_t b200_zerocheck_univariate_finish(b200_ctx *ctx, const b200_dev_ptr *multilins, const uint32_t *tower_levels,
										 uint32_t n_multilins, uint32_t n_vars, uint32_t skip_rounds, b200_dev_ptr eq_ind, uint64_t n_eq,
										 const b200_expr *const *compositions, const uint32_t *composition_degrees,
										 uint32_t n_compositions, uint32_t max_domain_size, b200_dev_ptr store, uint64_t store_elems,
										 uint64_t *host_round_evals);

#ifdef __cplusplus
}
#endif
#endif /* BINIUS_B200_H */
