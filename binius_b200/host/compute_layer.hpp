// compute_layer.hpp -- C++ host-side mirror of the reference's `binius_compute` trait surface over the
// C ABI (include/binius_b200.h).  The reference is compiled (Rust) code and its toolchain is absent
// from this image, so this header is the compiled-language statement of the plugin interface; names,
// argument meaning and error behaviour follow
//   ComputeLayer / ComputeLayerExecutor / KernelExecutor   crates/compute/src/layer.rs:22-88, 100-510, 518-590
//   KernelMemMap / KernelBuffer                             crates/compute/src/layer.rs:595-704
//   ComputeMemory (ALIGNMENT = 1) / SubfieldSlice           crates/compute/src/memory.rs:69-281
//   BumpAllocator                                           crates/compute/src/alloc.rs:31-105
//   ComputeHolder / ComputeData                             crates/compute/src/layer.rs:732-776
//   AdditiveNTT                                             crates/ntt/src/additive_ntt.rs:58-166
// Errors are C++ exceptions carrying the compute::Error / ntt::Error class.
#pragma once
#include <cstdint>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/binius_b200.h"

namespace binius_b200 {

struct F128 {
	uint64_t lo = 0, hi = 0;
	bool operator==(const F128 &o) const { return lo == o.lo && hi == o.hi; }
};

struct Error : std::runtime_error {
	int32_t code;
	Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};
struct InputValidation : Error { using Error::Error; };
struct AllocError : Error { using Error::Error; };
struct DeviceError : Error { using Error::Error; };
struct NttError : Error { using Error::Error; };

// FSlice / FSliceMut: (device pointer, length); any element offset is valid (ALIGNMENT = 1)
struct DevSlice {
	uint8_t *ptr = nullptr;
	uint64_t n = 0;
	static constexpr uint64_t ALIGNMENT = 1;
	uint64_t len() const { return n; }
	bool is_empty() const { return n == 0; }
	DevSlice slice(uint64_t start, uint64_t end) const {
		if (start > end || end > n) throw std::out_of_range("DevSlice::slice");
		return DevSlice{ptr + 16 * start, end - start};
	}
	std::pair<DevSlice, DevSlice> split_at(uint64_t mid) const { return {slice(0, mid), slice(mid, n)}; }
	std::pair<DevSlice, DevSlice> split_half() const { return split_at(n / 2); }
};
struct SubfieldSlice {
	DevSlice slice;
	uint32_t tower_level;
};
struct SlicesBatch {
	std::vector<DevSlice> rows;
	uint64_t row_len;
};

class BumpAllocator {
	DevSlice buf_;
	uint64_t off_ = 0;

  public:
	explicit BumpAllocator(DevSlice buffer) : buf_(buffer) {}
	DevSlice alloc(uint64_t n) {
		if (n > buf_.n - off_) throw AllocError(B200_ERR_ALLOC, "allocator is out of memory");
		DevSlice s = buf_.slice(off_, off_ + n);
		off_ += n;
		return s;
	}
	uint64_t remaining() const { return buf_.n - off_; }
	uint64_t capacity() const { return buf_.n; }
};

struct KernelMemMap {
	enum Kind { Chunked, ChunkedMut, Local } kind;
	DevSlice data;
	uint32_t log_min_chunk_size = 0, log_size = 0;
	static KernelMemMap chunked(DevSlice d, uint32_t log_min) { return {Chunked, d, log_min, 0}; }
	static KernelMemMap chunked_mut(DevSlice d, uint32_t log_min) { return {ChunkedMut, d, log_min, 0}; }
	static KernelMemMap local(uint32_t log_size) { return {Local, DevSlice{}, 0, log_size}; }
};
struct KernelBuffer {
	DevSlice data;
	bool is_mut;
	DevSlice to_ref() const { return data; }
	uint64_t len() const { return data.n; }
};
struct OpValue {
	uint32_t slot;
};

struct ExprStep : b200_expr_step {
	static ExprStep add(uint32_t l, uint32_t r) { return mk(0, l, r, {}); }
	static ExprStep mul(uint32_t l, uint32_t r) { return mk(1, l, r, {}); }
	static ExprStep pow(uint32_t l, uint64_t e) { return mk(2, l, e, {}); }
	static ExprStep constant(F128 c) { return mk(3, 0, 0, c); }
	static ExprStep var(uint32_t i) { return mk(4, i, 0, {}); }

  private:
	static ExprStep mk(uint32_t op, uint32_t l, uint64_t r, F128 c) {
		ExprStep s;
		s.op = op; s.l = l; s.r = r; s.c_lo = c.lo; s.c_hi = c.hi;
		return s;
	}
};

class B200Layer;

class ExprEval {
	b200_expr *h_ = nullptr;
	friend class B200Layer;
	friend class B200KernelExec;
	friend class B200Exec;

  public:
	ExprEval() = default;
	ExprEval(const ExprEval &) = delete;
	ExprEval(ExprEval &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
	~ExprEval() { if (h_) b200_expr_free(h_); }
	uint32_t n_vars() const { return b200_expr_n_vars(h_); }
	const b200_expr *raw() const { return h_; }
};

class B200KernelExec {
	B200Layer &l_;

  public:
	explicit B200KernelExec(B200Layer &l) : l_(l) {}
	OpValue decl_value(F128 init);
	void sum_composition_evals(const SlicesBatch &inputs, const ExprEval &composition, F128 batch_coeff, OpValue accumulator);
	void add(uint32_t log_len, DevSlice src1, DevSlice src2, DevSlice dst);
	void add_assign(uint32_t log_len, DevSlice src, DevSlice dst);
};

class B200Ntt;

class B200Exec {
	B200Layer &l_;
	friend class B200Layer;
	explicit B200Exec(B200Layer &l) : l_(l) {}

  public:
	using KernelFn = std::function<std::vector<OpValue>(B200KernelExec &, uint32_t /*log_chunks*/, std::vector<KernelBuffer> &)>;
	template <class A, class B> auto join(A op1, B op2) { auto r1 = op1(*this); auto r2 = op2(*this); return std::make_pair(r1, r2); }
	std::vector<OpValue> accumulate_kernels(const KernelFn &map, const std::vector<KernelMemMap> &mem_maps);
	void map_kernels(const KernelFn &map, const std::vector<KernelMemMap> &mem_maps) { accumulate_kernels(map, mem_maps); }
	OpValue inner_product(SubfieldSlice a_in, DevSlice b_in);
	void tensor_expand(uint32_t log_n, const std::vector<F128> &coordinates, DevSlice data);
	void fold_left(SubfieldSlice mat, DevSlice vec, DevSlice out);
	void fold_right(SubfieldSlice mat, DevSlice vec, DevSlice out);
	void fri_fold(const B200Ntt &ntt, uint32_t log_len, uint32_t log_batch_size, const std::vector<F128> &challenges, DevSlice data_in, DevSlice data_out);
	void extrapolate_line(DevSlice evals_0, DevSlice evals_1, F128 z);
	void compute_composite(const SlicesBatch &inputs, DevSlice output, const ExprEval &composition);
	void pairwise_product_reduce(DevSlice input, const std::vector<DevSlice> &round_outputs);
};

class B200Layer {
	b200_ctx *ctx_ = nullptr;
	std::recursive_mutex exec_mu_;
	std::vector<DevSlice> scratch_;
	friend class B200Exec;
	friend class B200KernelExec;
	friend class B200Ntt;

  public:
	explicit B200Layer(int device = 0) {
		int32_t rc = b200_ctx_create(device, &ctx_);
		if (rc) throw DeviceError(rc, "b200_ctx_create failed: no usable sm_100 GPU (no CPU fallback)");
	}
	B200Layer(const B200Layer &) = delete;
	~B200Layer() {
		for (auto &s : scratch_) b200_dev_free(ctx_, s.ptr);
		b200_ctx_destroy(ctx_);
	}
	b200_ctx *ctx() const { return ctx_; }
	void check(int32_t rc) const {
		if (!rc) return;
		std::string m = b200_last_error(ctx_);
		if (rc == B200_ERR_INPUT_VALIDATION) throw InputValidation(rc, m);
		if (rc == B200_ERR_ALLOC) throw AllocError(rc, m);
		if (rc >= 11 && rc <= 16) throw NttError(rc, m);
		throw DeviceError(rc, m);
	}
	DevSlice dev_alloc(uint64_t n) {
		void *p;
		check(b200_dev_alloc(ctx_, n, &p));
		return DevSlice{(uint8_t *)p, n};
	}
	void dev_free(DevSlice s) { check(b200_dev_free(ctx_, s.ptr)); }
	// ComputeLayer
	void copy_h2d(const F128 *src, uint64_t n, DevSlice dst) {
		if (n != dst.n) throw InputValidation(1, "src and dst must have the same length");
		check(b200_copy_h2d(ctx_, src, dst.ptr, n));
		check(b200_sync(ctx_));
	}
	void copy_d2h(DevSlice src, F128 *dst, uint64_t n) {
		if (n != src.n) throw InputValidation(1, "src and dst must have the same length");
		check(b200_copy_d2h(ctx_, src.ptr, dst, n));
	}
	void copy_d2d(DevSlice src, DevSlice dst) {
		if (src.n != dst.n) throw InputValidation(1, "src and dst must have the same length");
		check(b200_copy_d2d(ctx_, src.ptr, dst.ptr, src.n));
	}
	ExprEval compile_expr(const std::vector<ExprStep> &steps) {
		ExprEval e;
		check(b200_expr_compile(ctx_, steps.data(), (uint32_t)steps.size(), &e.h_));
		return e;
	}
	void fill(DevSlice s, F128 v) {
		uint64_t w[2] = {v.lo, v.hi};
		check(b200_fill(ctx_, s.ptr, s.n, w));
	}
	// an execute is a scope over the context's result slots: executes from several host threads take turns (every other
	// call is serialised by the context's own lock inside the library)
	std::vector<F128> execute(const std::function<std::vector<OpValue>(B200Exec &)> &f) {
		std::lock_guard<std::recursive_mutex> scope(exec_mu_);
		check(b200_results_reset(ctx_));
		B200Exec ex(*this);
		std::vector<OpValue> vals = f(ex);
		std::vector<uint32_t> slots;
		for (auto v : vals) slots.push_back(v.slot);
		std::vector<F128> out(vals.size());
		check(b200_results_fetch(ctx_, slots.data(), (uint32_t)slots.size(), (uint64_t *)out.data()));
		return out;
	}
};

// ---- out-of-line members ----------------------------------------------------------------------
inline OpValue B200KernelExec::decl_value(F128 init) {
	uint64_t w[2] = {init.lo, init.hi};
	OpValue v;
	l_.check(b200_kernel_decl_value(l_.ctx_, w, &v.slot));
	return v;
}
inline void B200KernelExec::sum_composition_evals(const SlicesBatch &in, const ExprEval &e, F128 c, OpValue acc) {
	std::vector<b200_dev_ptr> p;
	for (auto &r : in.rows) p.push_back(r.ptr);
	uint64_t w[2] = {c.lo, c.hi};
	l_.check(b200_kernel_sum_composition_evals(l_.ctx_, p.data(), (uint32_t)p.size(), in.row_len, e.h_, w, acc.slot));
}
inline void B200KernelExec::add(uint32_t log_len, DevSlice a, DevSlice b, DevSlice d) { l_.check(b200_kernel_add(l_.ctx_, log_len, a.ptr, b.ptr, d.ptr)); }
inline void B200KernelExec::add_assign(uint32_t log_len, DevSlice s, DevSlice d) { l_.check(b200_kernel_add_assign(l_.ctx_, log_len, s.ptr, d.ptr)); }

inline std::vector<OpValue> B200Exec::accumulate_kernels(const KernelFn &map, const std::vector<KernelMemMap> &mem_maps) {
	if (mem_maps.empty()) throw InputValidation(1, "Many variant must have at least one entry");
	// a kernel scope: the ops the closure issues are recorded by the library and lowered together at scope end
	// (the bivariate round-evaluation closure becomes tensor-core inner-product jobs; Locals are never written)
	l_.check(b200_kernel_scope_begin(l_.ctx_));
	std::vector<OpValue> out;
	try {
		std::vector<KernelBuffer> bufs;
		for (auto &m : mem_maps) {
			if (m.kind == KernelMemMap::Local) {
				void *p;
				l_.check(b200_kernel_local(l_.ctx_, m.log_size, &p));
				bufs.push_back({DevSlice{(uint8_t *)p, 1ull << m.log_size}, true});
			} else {
				bufs.push_back({m.data, m.kind == KernelMemMap::ChunkedMut});
			}
		}
		B200KernelExec kex(l_);
		out = map(kex, 0, bufs);  // the layer always selects log_chunks = 0 (layer.rs:149-160 allows any value in range)
	} catch (...) {
		b200_kernel_scope_end(l_.ctx_);
		throw;
	}
	l_.check(b200_kernel_scope_end(l_.ctx_));
	return out;
}
inline OpValue B200Exec::inner_product(SubfieldSlice a, DevSlice b) {
	OpValue v;
	l_.check(b200_inner_product(l_.ctx_, a.slice.ptr, a.slice.n, a.tower_level, b.ptr, b.n, &v.slot));
	return v;
}
inline void B200Exec::tensor_expand(uint32_t log_n, const std::vector<F128> &c, DevSlice data) {
	l_.check(b200_tensor_expand(l_.ctx_, data.ptr, data.n, log_n, (const uint64_t *)c.data(), (uint32_t)c.size()));
}
inline void B200Exec::fold_left(SubfieldSlice m, DevSlice v, DevSlice o) { l_.check(b200_fold_left(l_.ctx_, m.slice.ptr, m.slice.n, m.tower_level, v.ptr, v.n, o.ptr, o.n)); }
inline void B200Exec::fold_right(SubfieldSlice m, DevSlice v, DevSlice o) { l_.check(b200_fold_right(l_.ctx_, m.slice.ptr, m.slice.n, m.tower_level, v.ptr, v.n, o.ptr, o.n)); }
inline void B200Exec::extrapolate_line(DevSlice e0, DevSlice e1, F128 z) {
	uint64_t w[2] = {z.lo, z.hi};
	l_.check(b200_extrapolate_line(l_.ctx_, e0.ptr, e0.n, e1.ptr, e1.n, w));
}
inline void B200Exec::compute_composite(const SlicesBatch &in, DevSlice out, const ExprEval &e) {
	if (e.n_vars() != in.rows.size()) throw InputValidation(1, "composition not match with input");
	std::vector<b200_dev_ptr> p;
	for (auto &r : in.rows) p.push_back(r.ptr);
	l_.check(b200_compute_composite(l_.ctx_, p.data(), (uint32_t)p.size(), in.row_len, out.ptr, out.n, e.h_));
}
inline void B200Exec::pairwise_product_reduce(DevSlice in, const std::vector<DevSlice> &outs) {
	std::vector<b200_dev_ptr> p;
	std::vector<uint64_t> n;
	for (auto &o : outs) { p.push_back(o.ptr); n.push_back(o.n); }
	l_.check(b200_pairwise_product_reduce(l_.ctx_, in.ptr, in.n, p.data(), n.data(), (uint32_t)outs.size()));
}

struct NTTShape {
	uint32_t log_x = 0, log_y = 0, log_z = 0;
};

// AdditiveNTT<F> for F = BinaryField{8,16,32}b (field_log_bits 3,4,5); data = host `&mut [P]`
class B200Ntt {
	B200Layer &l_;
	b200_ntt *h_ = nullptr;
	friend class B200Exec;

  public:
	B200Ntt(B200Layer &l, uint32_t field_log_bits, uint32_t log_domain_size) : l_(l) { l_.check(b200_ntt_create(l_.ctx_, field_log_bits, log_domain_size, &h_)); }
	B200Ntt(const B200Ntt &) = delete;
	~B200Ntt() { b200_ntt_destroy(h_); }
	uint32_t log_domain_size() const { return b200_ntt_log_domain_size(h_); }
	const b200_ntt *raw() const { return h_; }
	F128 get_subspace_eval(uint32_t i, uint64_t j) const {
		F128 o;
		if (b200_ntt_get_subspace_eval(h_, i, j, &o.lo)) throw InputValidation(1, "get_subspace_eval out of range");
		return o;
	}
	void forward_transform(void *data, uint32_t elem_log_bits, uint64_t n_elems, NTTShape s, uint64_t coset, uint32_t coset_bits, uint32_t skip_rounds) const {
		l_.check(b200_ntt_forward_host(l_.ctx_, h_, data, elem_log_bits, n_elems, s.log_x, s.log_y, s.log_z, coset, coset_bits, skip_rounds));
	}
	void inverse_transform(void *data, uint32_t elem_log_bits, uint64_t n_elems, NTTShape s, uint64_t coset, uint32_t coset_bits, uint32_t skip_rounds) const {
		l_.check(b200_ntt_inverse_host(l_.ctx_, h_, data, elem_log_bits, n_elems, s.log_x, s.log_y, s.log_z, coset, coset_bits, skip_rounds));
	}
};
inline void B200Exec::fri_fold(const B200Ntt &ntt, uint32_t log_len, uint32_t log_batch, const std::vector<F128> &ch, DevSlice in, DevSlice out) {
	l_.check(b200_fri_fold(l_.ctx_, ntt.h_, log_len, log_batch, (const uint64_t *)ch.data(), (uint32_t)ch.size(), in.ptr, in.n, out.ptr, out.n));
}

// ComputeHolder: host arena + device arena + the layer (cf. FastCpuLayerHolder::new, examples/keccak.rs:119-122)
struct ComputeData {
	B200Layer &hal;
	std::vector<F128> &host_mem;
	BumpAllocator dev_alloc;
};
class B200LayerHolder {
	B200Layer layer_;
	std::vector<F128> host_mem_;
	DevSlice dev_mem_;

  public:
	B200LayerHolder(uint64_t host_mem_size, uint64_t dev_mem_size, int device = 0) : layer_(device), host_mem_(host_mem_size) {
		dev_mem_ = layer_.dev_alloc(dev_mem_size);
		layer_.fill(dev_mem_, F128{});
	}
	~B200LayerHolder() { b200_dev_free(layer_.ctx(), dev_mem_.ptr); }
	ComputeData to_data() { return ComputeData{layer_, host_mem_, BumpAllocator(dev_mem_)}; }
};

}  // namespace binius_b200
