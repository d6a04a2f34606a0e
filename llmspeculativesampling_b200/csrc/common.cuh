// Shared device helpers for the sm_100a speculative-decoding kernels.
//
// Everything here is written for B200 (sm_100a): 1-D TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx) stage vocabulary slices in shared memory, thread-block clusters + distributed shared
// memory split one vocabulary row over up to 8 CTAs, and all selection / sampling decisions are
// made with order-preserving integer keys and 64-bit fixed-point weights so that results do not
// depend on the order of parallel reductions.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace sd {

constexpr int kScaleBits = 40;      // fixed-point fraction bits of sampling weights (oracle/ref_ops.py)
constexpr int kUBits = 24;          // uniforms are consumed as 24-bit integers
constexpr float kProbGuard = 1e-9f; // reference sampling/utils.py:228

// error bits written (atomicOr) into the caller's device err_flag
enum : int {
  kErrNanLogit = 1,   // NaN or +inf logit / non-finite probability  -> RuntimeError('norm logits error')
  kErrEmptyRow = 2,   // no positive weight to sample from            -> RuntimeError('prob error')
  kErrZeroQ = 4,      // draft probability of a drafted token is 0    -> RuntimeError('s')
  kErrBadToken = 8,   // drafted token id outside [0, V)
};

enum : int { kF32 = 0, kBF16 = 1, kF16 = 2 };

// ----------------------------------------------------------------------------------------------
// order-preserving float <-> uint32 key (larger float  <=>  larger key; -0 < +0 is harmless here)
__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}
// the float `ulps` representable steps below f (saturating at -inf's key)
__device__ __forceinline__ float float_down(float f, uint32_t ulps) {
  uint32_t k = f2key(f);
  const uint32_t kmin = 0x007fffffu;  // key(-inf)
  k = (k > kmin + ulps) ? (k - ulps) : kmin;
  return key2f(k);
}

// ----------------------------------------------------------------------------------------------
// element access on a staged slice (raw storage dtype -> fp32)
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kPerVec = 4;   // elements per 16-byte vector
  static constexpr uint32_t kNegInfWord = 0xff800000u;   // a 32-bit word of -inf elements
  __device__ static __forceinline__ void unpack(const uint4& v, float (&o)[4]) {
    o[0] = __uint_as_float(v.x); o[1] = __uint_as_float(v.y);
    o[2] = __uint_as_float(v.z); o[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ float to_f(float x) { return x; }
  __device__ static __forceinline__ float neg_inf() { return -INFINITY; }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kPerVec = 8;
  static constexpr uint32_t kNegInfWord = 0xff80ff80u;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&o)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[2 * i] = __uint_as_float(w[i] << 16);
      o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
  __device__ static __forceinline__ __nv_bfloat16 neg_inf() { return __float2bfloat16(-INFINITY); }
};
template <> struct Elem<__half> {
  static constexpr int kPerVec = 8;
  static constexpr uint32_t kNegInfWord = 0xfc00fc00u;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&o)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 f = __half22float2(h);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
  __device__ static __forceinline__ float to_f(__half x) { return __half2float(x); }
  __device__ static __forceinline__ __half neg_inf() { return __float2half(-INFINITY); }
};

// NaN-propagating maximum of (m, all elements of one 16-byte vector).  16-bit types are reduced with packed
// two-lane instructions before a single conversion to fp32 (the conversion is monotone, so the result is exact).
template <typename T> __device__ __forceinline__ float vec_max_nan(float m, const uint4& v);
template <> __device__ __forceinline__ float vec_max_nan<float>(float m, const uint4& v) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)));
  asm("max.NaN.f32 %0, %0, %1;" : "+f"(d) : "f"(__uint_as_float(v.z)));
  asm("max.NaN.f32 %0, %0, %1;" : "+f"(d) : "f"(__uint_as_float(v.w)));
  asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(d));
  return m;
}
template <> __device__ __forceinline__ float vec_max_nan<__nv_bfloat16>(float m, const uint4& v) {
  uint32_t a, b;
  asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(a) : "r"(v.x), "r"(v.y));
  asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(b) : "r"(v.z), "r"(v.w));
  asm("max.NaN.bf16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(__uint_as_float(a << 16)), "f"(__uint_as_float(a & 0xffff0000u)));
  asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(d));
  return m;
}
template <> __device__ __forceinline__ float vec_max_nan<__half>(float m, const uint4& v) {
  uint32_t a, b;
  asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(a) : "r"(v.x), "r"(v.y));
  asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(b) : "r"(v.z), "r"(v.w));
  asm("max.NaN.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&a));
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(f.x), "f"(f.y));
  asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(d));
  return m;
}

// ----------------------------------------------------------------------------------------------
// warp / block reductions
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive scan over the 32 lanes
__device__ __forceinline__ unsigned long long warp_scan_incl(unsigned long long v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ int warp_scan_incl(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// bitonic sort of one value per lane, descending (lane 0 = largest)
__device__ __forceinline__ float warp_sort_desc(float v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      float o = __shfl_xor_sync(0xffffffffu, v, j);
      bool up = (lane & k) == 0;       // block sorted descending when `up`
      bool lower = (lane & j) == 0;
      v = (lower == up) ? fmaxf(v, o) : fminf(v, o);
    }
  }
  return v;
}

// ----------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (global -> shared::cta), sm_90+/sm_100a
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// an mbarrier must be invalidated before its memory is re-initialised or re-purposed
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// suspend-time hint of every mbarrier wait: a waiting warp is parked by the hardware (it wakes as soon as the phase
// completes) instead of re-issuing try_wait — spinning warps were measured to take the issue slots of the warps that
// share their scheduler (a 4-warp sort slowed down 4x next to 12 spinning warps)
constexpr uint32_t kMbarSuspendNs = 20000u;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase), "r"(kMbarSuspendNs)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {
  }
}
// bulk async copy global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// bulk async copy shared::cta -> global (bulk-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// streaming (evict-first) 16-byte global store / load
__device__ __forceinline__ void st_cs_v4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 32-byte streaming store (sm_100: STG.E.256; the address must be 32-byte aligned): a thread that owns 8 consecutive
// floats writes ONE full 32-byte sector instead of two half sectors from two instructions
__device__ __forceinline__ void st_cs_v8(float* p, const float* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
// the PV fp32 results of one input vector: one 32-byte store for 16-bit inputs when the destination is 32-byte aligned
// (two half-sector stores from two instructions were measured to cost the dense bf16 path 20 %), 16-byte stores otherwise
template <int PV>
__device__ __forceinline__ void st_cs_vec(float* dst, const float* v) {
  if (PV == 8 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    st_cs_v8(dst, v);
  } else {
#pragma unroll
    for (int j = 0; j < PV; j += 4) st_cs_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ----------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 — one issue slot for two fp32 lanes, every IEEE rounding mode)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 fma2_rd(f32x2 a, f32x2 b, f32x2 c) {        // round towards -inf
  f32x2 d;
  asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// (CTA scheduling, shared-memory carve-up, barrier initialisation) before its predecessor in the stream has finished;
// it must execute pdl_wait() before its first global-memory access.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// fixed-point sampling weights (oracle/ref_ops.py: sampling_weights / icdf_sample)
//   w = floor(p * 2^(40 - e)),  e = frexp exponent of the row maximum
__device__ __forceinline__ int frexp_exp(float mx) {
  int e;
  (void)frexpf(mx, &e);
  return e;
}
__device__ __forceinline__ unsigned long long weight_of(float p, int e) {
  // ldexpf by a power of two is exact; p <= max < 2^e so the product is < 2^40
  return (p > 0.f) ? __float2ull_rd(ldexpf(p, kScaleBits - e)) : 0ull;
}
__device__ __forceinline__ uint32_t u_to_int(float u) {
  float s = floorf(u * 16777216.0f);
  s = fminf(fmaxf(s, 0.f), 16777215.0f);
  return static_cast<uint32_t>(s);
}
// t = (total * m) >> 24 without 128-bit overflow
__device__ __forceinline__ unsigned long long scale_target(unsigned long long total, uint32_t m) {
  unsigned long long hi = total >> kUBits, lo = total & ((1ull << kUBits) - 1ull);
  return hi * m + ((lo * m) >> kUBits);
}

}  // namespace sd
