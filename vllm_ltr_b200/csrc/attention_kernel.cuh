// attention_kernel.cuh — single-query (decode) attention over the paged KV cache for sm_100a.
//
// Replaces paged_attention_kernel / _v1_kernel / _v2_kernel / _v2_reduce_kernel of the
// reference (csrc/attention/attention_kernels.cu:97-602).  Same inputs, same paged layouts
// (K [blk, kvh, D/x, bs, x], V [blk, kvh, D, bs]), same masking / ALiBi / 1e-6 denominator
// semantics; a different machine mapping:
//
//   reference                                   this kernel
//   ----------------------------------------    ---------------------------------------------
//   grid (q_head, seq, partition): every q      work item = (seq, partition, kv_head): one warp
//   head of a GQA group re-reads the same KV    streams the KV of a kv head ONCE and applies it
//                                               to all G = Hq/Hkv query heads of the group
//   8-byte LDG of K straight from global,       each (block, kv_head) K tile and V tile is one
//   no staging, no async copy                   contiguous span: moved HBM->smem by TMA bulk
//                                               copies (cp.async.bulk, SASS UBLKCP) into a
//                                               per-warp mbarrier ring that runs ACROSS work
//                                               items (the next item's tiles are already in
//                                               flight while the current one finishes)
//   logits for the whole context in smem        online softmax per KV block in registers:
//   (4*ctx bytes -> 8K..57K token limit),       no context-length limit, no CTA-wide barrier
//   3 __syncthreads + smem tree per CTA         in the main loop (warps are independent)
//   q.k with cvt f16->f32 + FFMA; p.v in        G == 1: FHFMA (fma.rn.f32.f16/.bf16, sm_100
//   half precision (mul.f16x2)                  mixed-precision FMA), fp32 accumulation.
//                                               G >= 2: tensor cores, both products transposed
//                                               (S^T = K.Q^T, O^T = V^T.P^T): tokens / head-dim
//                                               rows fill the 16-row M of mma.sync.m16n8k16, the
//                                               query heads of the kv head ride its 8-wide N; K / V
//                                               fragments by ldmatrix straight from the paged tile
//                                               layout, P^T by movmatrix from the S^T accumulators
//   static grid; v2 = second kernel that        persistent grid; contexts are cut into
//   merges 512-token partitions through         partitions sized to fill the machine, a seq's
//   global memory                               partitions are merged by whichever warp finishes
//                                               LAST (atomic ticket) - one launch, no tail
//
// Bound: HBM.  Algorithmic bytes per call = sum_i L_i*2*Hkv*D*s_cache + q + out + tables.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace vltr {
namespace attn {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxPlanSeqs = 1024;   // per-CTA partition plan lives in shared memory up to here
constexpr int kPlanBytes = 8192;     // cum[n+1] ints + 4 x 66 + 8 ints (tail histograms) + n uint16

struct Params {
  void* out;                    // [num_seqs, num_heads, D]  T
  const void* q;                // [num_seqs, num_heads, D]  T, row stride q_stride
  const void* k_cache;
  const void* v_cache;
  const int32_t* block_tables;  // [num_seqs, max_blocks_per_seq]
  const int32_t* context_lens;  // [num_seqs]
  const float* alibi_slopes;    // [num_heads] or null
  // partial results of partitioned contexts
  float* exp_sums;              // [num_seqs, num_heads, n_part]
  float* max_logits;            // [num_seqs, num_heads, n_part]
  void* tmp_out;                // [num_seqs, num_heads, n_part, D]  T (v2 caller scratch) or float
  int always_partial;           // v2: fill the caller's partials even for single-partition seqs
  int tmp_is_f32;
  int num_seqs, num_heads, num_kv_heads, group, passes;
  int max_blocks_per_seq;
  int n_part;                   // partitions of the longest context (second-to-last tmp dim)
  int part_blocks;              // KV blocks per partition (<= 64); the MINIMUM when `adaptive`
  int adaptive;                 // v1 + plan: the kernel sizes partitions from the actual contexts
  int part_blocks_cap;          // adaptive: largest partition (KV blocks)
  float scale;
  int64_t q_stride;             // elements
  int64_t kv_block_stride;      // bytes
  int64_t kv_head_stride;       // bytes
  int stages;
  int use_plan;                 // 1: items enumerate only non-empty partitions (num_seqs <= kMaxPlanSeqs)
  int* counters;                // [0] next item, [1] finished warps   (self-resetting)
  int* tickets;                 // [num_seqs * passes * num_kv_heads] partition tickets, all zero
  long long* trace;             // development: 16 stamps per warp (tools/attn_tune.py) or null
  int defer_merge;              // 1: partitions are merged by merge_partials_kernel (second launch)
  int* plan_out;                // defer_merge: [0] = KV blocks per partition the kernel chose
};
// per-warp timeline for tools/attn_tune.py --trace: build with `make EXTRA=-DVLTR_ATTN_TRACE=1`
// (kept out of the production build: the extra live state costs registers in the tile loop)
#ifndef VLTR_ATTN_TRACE
#define VLTR_ATTN_TRACE 0
#endif
// KV blocks per loop iteration of the tensor-core path: 0 = one, 1 = two for e5m2 caches,
// 2 = two for 2-byte caches as well.  Measured (profiles/r02_attn_pair.txt): two blocks per
// iteration change nothing (e5m2 trace batch 47.5 vs 46.3 us, bf16 GQA 77.7 vs 76.1 us) - the
// tile loop is not what bounds these launches - so the default stays one.
#ifndef VLTR_ATTN_PAIR
#define VLTR_ATTN_PAIR 0
#endif
// two interleaved accumulator chains for S^T = K.Q^T (tensor-core path)
#ifndef VLTR_ATTN_SPLIT_QK
#define VLTR_ATTN_SPLIT_QK 1
#endif
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t lds4(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// mixed-precision FMA (sm_100): d = a.f16 * b.f16 + c.f32, both halves of a packed word
template <typename T>
__device__ __forceinline__ float fhfma2(uint32_t a, uint32_t b, float c);
template <>
__device__ __forceinline__ float fhfma2<__half>(uint32_t a, uint32_t b, float c) {
  float r;
  asm("{.reg .b16 al,ah,bl,bh; .reg .f32 t;\n"
      "mov.b32 {al,ah}, %1; mov.b32 {bl,bh}, %2;\n"
      "fma.rn.f32.f16 t, al, bl, %3;\n"
      "fma.rn.f32.f16 %0, ah, bh, t;}"
      : "=f"(r)
      : "r"(a), "r"(b), "f"(c));
  return r;
}
template <>
__device__ __forceinline__ float fhfma2<__nv_bfloat16>(uint32_t a, uint32_t b, float c) {
  float r;
  asm("{.reg .b16 al,ah,bl,bh; .reg .f32 t;\n"
      "mov.b32 {al,ah}, %1; mov.b32 {bl,bh}, %2;\n"
      "fma.rn.f32.bf16 t, al, bl, %3;\n"
      "fma.rn.f32.bf16 %0, ah, bh, t;}"
      : "=f"(r)
      : "r"(a), "r"(b), "f"(c));
  return r;
}

// D(16x8, f32) += A(16x16) * B(16x8); MT = __half or __nv_bfloat16 operands
template <typename MT>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  if constexpr (std::is_same<MT, __half>::value) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

// 2^x on the SFU, flush-to-zero: one MUFU.EX2.  exp2f() wraps the same instruction in a range
// fix-up for results below 2^-126 (compare, halve, square: 3 more instructions per call); softmax
// weights that small round to zero in the f16 / bf16 operand anyway.
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <typename MT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (std::is_same<MT, __half>::value) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}
template <typename MT>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (std::is_same<MT, __half>::value)
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
  else
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

// ---- compile-time geometry ------------------------------------------------------------------------
// GT: query heads handled per work item ("pass").  MMA: tensor-core path (GT = 8 or 16 rows).
template <typename T, typename CT, int D_, int BS_, int GT_, bool MMA_>
struct Cfg {
  static constexpr int D = D_, BS = BS_, GT = GT_;
  static constexpr bool MMA = MMA_;
  static constexpr bool FP8 = !std::is_same<T, CT>::value;
  static constexpr int EPC = 16 / sizeof(CT);       // elements per 16-byte cache chunk (== x)
  static constexpr int NCH = D / EPC;               // K chunks per token
  static constexpr int TG = 32 / BS;                // lanes sharing one token in q.k (FMA path)
  static constexpr int KJ = (NCH + TG - 1) / TG;    // K chunks per lane
  static constexpr int LPR = BS / EPC;              // lanes sharing one V row
  static constexpr int RPI = 32 / (LPR > 0 ? LPR : 1);  // V rows per iteration
  static constexpr int RI = (D + RPI - 1) / RPI;    // V row iterations per lane
  static constexpr bool HALF = std::is_same<T, __half>::value;
  static constexpr bool PACKED = HALF || (std::is_same<T, __nv_bfloat16>::value && !FP8);
  using QS = typename std::conditional<PACKED, T, float>::type;   // q / p storage in smem (FMA)
  // tensor-core operand type: the cache type for 2-byte caches, f16 for e5m2 caches (e5m2 -> f16
  // is a byte shift; a bf16 query converts to f16 exactly in the normal range)
  using MT = typename std::conditional<FP8, __half, T>::type;
  static constexpr int QW = EPC * sizeof(QS) / 16;  // 16-byte words of q (or p) per cache chunk
  static constexpr int TILE_BYTES = D * BS * (int)sizeof(CT);
  static constexpr int Q_BYTES = MMA ? 0 : GT * D * (int)sizeof(QS);
  static constexpr int P_BYTES = MMA ? 0 : GT * BS * (int)sizeof(QS);
  static constexpr bool P_IN_REGS = GT * QW <= 8;
  static constexpr int NT = BS / 8;                 // MMA: n-tiles (8 tokens) per KV block
  static constexpr int KS = D / 16;                 // MMA: k-steps of q.k
  static constexpr int ND = D / 8;                  // MMA: n-tiles of p.v
  static constexpr int KT = (BS + 15) / 16;         // MMA: k-steps of p.v
  static constexpr bool R16 = GT > 8;               // MMA: rows 8..15 of the tile are live
  static_assert(D % EPC == 0 && 32 % BS == 0, "geometry");
  static_assert(!MMA || (D % 16 == 0 && (GT == 8 || GT == 16) && sizeof(T) == 2), "mma geometry");
  static_assert(MMA || (BS % EPC == 0), "fma path needs whole 16-byte chunks per V row");
};

template <typename C>
__host__ __device__ constexpr int warp_smem_bytes(int stages) {
  return stages * 2 * C::TILE_BYTES + ((C::Q_BYTES + 15) / 16) * 16 + ((C::P_BYTES + 15) / 16) * 16 +
         ((stages * 8 + 15) / 16) * 16;
}

// acc += dot(one 16-byte cache chunk, the matching q (or p) words)            (FMA path)
template <typename T, typename C>
__device__ __forceinline__ float dot_chunk(float acc, const uint4& kv, const uint4* qw) {
  const uint32_t k[4] = {kv.x, kv.y, kv.z, kv.w};
  if constexpr (C::PACKED && !C::FP8) {
    const uint32_t q[4] = {qw[0].x, qw[0].y, qw[0].z, qw[0].w};
#pragma unroll
    for (int w = 0; w < 4; ++w) acc = fhfma2<T>(k[w], q[w], acc);
  } else if constexpr (!C::FP8) {  // float cache, float q
    const uint32_t q[4] = {qw[0].x, qw[0].y, qw[0].z, qw[0].w};
#pragma unroll
    for (int w = 0; w < 4; ++w) acc = fmaf(__uint_as_float(k[w]), __uint_as_float(q[w]), acc);
  } else if constexpr (C::PACKED) {  // e5m2 cache, half q: e5m2 byte b == half bits (b << 8)
    const uint32_t q[8] = {qw[0].x, qw[0].y, qw[0].z, qw[0].w, qw[1].x, qw[1].y, qw[1].z, qw[1].w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      acc = fhfma2<__half>(__byte_perm(k[w], 0, 0x1404), q[2 * w], acc);
      acc = fhfma2<__half>(__byte_perm(k[w], 0, 0x3424), q[2 * w + 1], acc);
    }
  } else {  // e5m2 cache, float q (bf16 / float models): fp8 -> half -> float is exact
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t q[4] = {qw[w].x, qw[w].y, qw[w].z, qw[w].w};
      const uint32_t h01 = __byte_perm(k[w], 0, 0x1404), h23 = __byte_perm(k[w], 0, 0x3424);
      const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
      const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
      acc = fmaf(f01.x, __uint_as_float(q[0]), acc);
      acc = fmaf(f01.y, __uint_as_float(q[1]), acc);
      acc = fmaf(f23.x, __uint_as_float(q[2]), acc);
      acc = fmaf(f23.y, __uint_as_float(q[3]), acc);
    }
  }
  return acc;
}

// keep the first `valid` elements of a 16-byte cache chunk, zero the rest (tail of the last block:
// the reference zeroes out-of-context V explicitly because it may hold NaNs, :367-376)
template <typename C>
__device__ __forceinline__ void mask_chunk(uint4& v, int valid) {
  constexpr int EPW = C::EPC / 4;  // elements per 32-bit word
  constexpr int BITS = 32 / EPW;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = valid - i * EPW;
    n = n < 0 ? 0 : (n > EPW ? EPW : n);
    const uint32_t m = n == EPW ? 0xffffffffu : ((1u << (n * BITS)) - 1u);
    w[i] &= m;
  }
}

template <typename QS>
__device__ __forceinline__ QS to_qs(float x);
template <>
__device__ __forceinline__ float to_qs<float>(float x) { return x; }
template <>
__device__ __forceinline__ __half to_qs<__half>(float x) { return __float2half_rn(x); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_qs<__nv_bfloat16>(float x) {
  return __float2bfloat16_rn(x);
}
template <typename QS>
__device__ __forceinline__ float qs_to_f(QS x);
template <>
__device__ __forceinline__ float qs_to_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ float qs_to_f<__half>(__half x) { return __half2float(x); }
template <>
__device__ __forceinline__ float qs_to_f<__nv_bfloat16>(__nv_bfloat16 x) {
  return __bfloat162float(x);
}

// L2 (cache-global) load of one T as float: partials written by other SMs in this launch
template <typename T>
__device__ __forceinline__ float load_cg_f(const T* ptr) {
  if constexpr (sizeof(T) == 4) {
    return __ldcg(reinterpret_cast<const float*>(ptr));
  } else {
    const unsigned short raw = __ldcg(reinterpret_cast<const unsigned short*>(ptr));
    if constexpr (std::is_same<T, __half>::value)
      return __half2float(__ushort_as_half(raw));
    else
      return __bfloat162float(__ushort_as_bfloat16(raw));
  }
}

// ---- merge of a sequence's partitions (reference: paged_attention_v2_reduce_kernel, :511-602) ---
// out[h] = sum_j w_j * tmp[h][j],  w_j = exp_sum_j * exp(max_j - max) / (sum + 1e-6).
// One warp; lanes own 4 consecutive output elements per 128; HB heads x JB partitions of
// independent 8/16-byte L2 loads are kept in flight.
template <typename T>
__device__ __forceinline__ float4 load_cg_4(const void* base, int64_t elem, bool f32) {
  if (f32) return __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem));
  if constexpr (sizeof(T) == 2) {
    const uint2 raw = __ldcg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(base) + elem));
    float2 a, b;
    if constexpr (std::is_same<T, __half>::value) {
      a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
      b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    } else {
      a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    }
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem));
  }
}

// Merge of a sequence's partitions by one warp.  Every load here is an L2 round trip on the
// critical path of the warp that finished the sequence's last partition (the partials were just
// written by other SMs), so the loads are issued in as few dependent waves as possible: lane j
// fetches (max, sum) of partition j for HB heads at once, the first JB partial outputs of those
// heads are requested in the same wave (they do not depend on the weights), the weights are
// formed with shuffles while everything is in flight.  Up to 32 partitions per wave of
// statistics (the host caps a sequence at 16, the reference's v2 entry at max_ctx / 512).
template <typename T, int D>
__device__ __forceinline__ void merge_partials(const Params& p, int seq, int head0, int nheads,
                                               int np, int lane) {
  constexpr int HB = 4, JB = 4;
  constexpr int NV = (D + 127) / 128;   // 4-element vectors per lane
  for (int h0 = 0; h0 < nheads; h0 += HB) {
    int64_t base[HB];
    float mx[HB], inv[HB], wl[HB];      // wl: weight of partition `lane` (np <= 32)
#pragma unroll
    for (int hh = 0; hh < HB; ++hh) {
      const int h = min(h0 + hh, nheads - 1);
      base[hh] = ((int64_t)seq * p.num_heads + head0 + h) * p.n_part;
    }
    // wave 1: statistics of partitions lane, lane + 32, ... (one load pair per head when np <= 32)
    float mj[HB], sj[HB];
#pragma unroll
    for (int hh = 0; hh < HB; ++hh) {
      mj[hh] = lane < np ? __ldcg(p.max_logits + base[hh] + lane) : -INFINITY;
      sj[hh] = lane < np ? __ldcg(p.exp_sums + base[hh] + lane) : 0.f;
    }
    // wave 1, continued: the first JB partial outputs of every head
    float4 v[HB][JB][NV];
    auto fetch = [&](int j0) {
#pragma unroll
      for (int hh = 0; hh < HB; ++hh)
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const int j = min(j0 + jj, np - 1);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int d = 4 * lane + 128 * i;
            v[hh][jj][i] = (D % 128 == 0 || d < D)
                               ? load_cg_4<T>(p.tmp_out, (base[hh] + j) * D + d, p.tmp_is_f32 != 0)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
    };
    fetch(0);
#pragma unroll
    for (int hh = 0; hh < HB; ++hh) {
      float m = mj[hh];
      for (int j = lane + 32; j < np; j += 32) m = fmaxf(m, __ldcg(p.max_logits + base[hh] + j));
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = sj[hh] * expf(mj[hh] - m);          // lane >= np: 0 * exp(-inf) = 0
      wl[hh] = sum;
      for (int j = lane + 32; j < np; j += 32)
        sum += __ldcg(p.exp_sums + base[hh] + j) * expf(__ldcg(p.max_logits + base[hh] + j) - m);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      mx[hh] = m;
      inv[hh] = __fdividef(1.0f, sum + 1e-6f);
      wl[hh] *= inv[hh];
    }
    float4 a[HB][NV];
#pragma unroll
    for (int hh = 0; hh < HB; ++hh)
#pragma unroll
      for (int i = 0; i < NV; ++i) a[hh][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < np; j0 += JB) {
      if (j0 > 0) fetch(j0);
      float w[HB][JB];
#pragma unroll
      for (int hh = 0; hh < HB; ++hh)
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const int j = j0 + jj;
          float wj;
          if (j0 + JB <= 32) {          // warp-uniform: the weights of this chunk sit in lanes
            wj = __shfl_sync(0xffffffffu, wl[hh], j & 31);
          } else {
            const int jc = min(j, np - 1);
            wj = __ldcg(p.exp_sums + base[hh] + jc) *
                 expf(__ldcg(p.max_logits + base[hh] + jc) - mx[hh]) * inv[hh];
          }
          w[hh][jj] = j < np ? wj : 0.f;
        }
#pragma unroll
      for (int hh = 0; hh < HB; ++hh)
#pragma unroll
        for (int jj = 0; jj < JB; ++jj)
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            a[hh][i].x += v[hh][jj][i].x * w[hh][jj];
            a[hh][i].y += v[hh][jj][i].y * w[hh][jj];
            a[hh][i].z += v[hh][jj][i].z * w[hh][jj];
            a[hh][i].w += v[hh][jj][i].w * w[hh][jj];
          }
    }
#pragma unroll
    for (int hh = 0; hh < HB; ++hh) {
      if (h0 + hh < nheads) {
        T* o = reinterpret_cast<T*>(p.out) + ((int64_t)seq * p.num_heads + head0 + h0 + hh) * D;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int d = 4 * lane + 128 * i;
          if (D % 128 == 0 || d < D) {
            o[d] = ST<T>::from_f(a[hh][i].x);
            o[d + 1] = ST<T>::from_f(a[hh][i].y);
            o[d + 2] = ST<T>::from_f(a[hh][i].z);
            o[d + 3] = ST<T>::from_f(a[hh][i].w);
          }
        }
      }
    }
  }
}

// ---- work items -----------------------------------------------------------------------------------
// One item = one partition of one sequence for one kv head and one pass over its query heads.
struct Item {
  int valid;            // 0: queue exhausted
  int seq, part, pass, kvh;
  int ctx, nparts;      // context length and number of partitions of this sequence
  int blk0, ntiles;     // first KV block and number of KV blocks of this partition
  int issued;           // tiles handed to the TMA ring so far
  int bt0, bt1;         // lane i holds the block ids of tiles i and 32+i
};

// cum[s] = number of partitions of sequences < s (plan mode), in shared memory
// The queue ticket is fetched one item AHEAD (lane 0 keeps the pending id in a register and only
// broadcasts it here), so the atomic's round trip to L2 overlaps the previous item's work.
// The FIRST item of every warp is its own index (no atomic: one L2 round trip less in front of
// the first KV request); the queue counter hands out the items after those.
__device__ __forceinline__ int fetch_id(const Params& p, int lane, int first_dynamic) {
  int id = 0;
  if (lane == 0) id = atomicAdd(&p.counters[0], 1) + first_dynamic;
  return id;
}
// Order of the partial last partitions ("tails") in the queue, after all full partitions: bucket
// 65 first, then 64 ... 0.  A single-partition sequence sorts by its length in KV blocks (longest
// first).  The tail of a MULTI-partition sequence goes to bucket 65 whatever its length: whoever
// finishes a sequence's last partition also merges its partitions, and that merge should happen
// while the machine is still streaming, not after the queue has drained.
__device__ __forceinline__ int tail_bucket(int ctx, int rem, int part_tokens, int BS) {
  return ctx >= part_tokens ? 65 : (rem + BS - 1) / BS;
}
struct Plan {
  const int* cum;                   // [n+1] exclusive prefix sums of FULL partitions per sequence
  const unsigned short* tail_order; // sequences that end in a partial partition, longest tail first
  int n_full;                       // total number of full partitions
};
__device__ __forceinline__ void claim(Item& it, int pending_id, const Params& p, const Plan& plan,
                                      int total_items, int part_blocks, int BS, int lane) {
  const int* cum = plan.cum;
  const int part_tokens = part_blocks * BS;
  const int item = __shfl_sync(0xffffffffu, pending_id, 0);
  it.valid = item < total_items;
  it.issued = 0;
  it.ntiles = 0;
  if (!it.valid) return;
  it.kvh = item % p.num_kv_heads;
  int t = item / p.num_kv_heads;
  it.pass = t % p.passes;
  t /= p.passes;
  bool tail_item = false;
  if (p.use_plan) {
    // longest-processing-time-first order: all full partitions (part_tokens each), then the
    // partial tails sorted by length
    if (t < plan.n_full) {
      int lo = 0, hi = p.num_seqs;     // largest s with cum[s] <= t
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] <= t) lo = mid; else hi = mid;
      }
      it.seq = lo;
      it.part = t - cum[lo];
    } else {
      it.seq = plan.tail_order[t - plan.n_full];
      it.part = 0;
      tail_item = true;
    }
  } else {
    it.part = t % p.n_part;
    it.seq = t / p.n_part;
  }
  if (tail_item) {    // the tail is the sequence's last partition: its index needs the context
    it.ctx = p.context_lens[it.seq];
    it.part = it.ctx / part_tokens;
  }
  it.blk0 = it.part * part_blocks;
  // block ids and the context length are independent loads (one L2 round trip, not two): the
  // table row is read up to its allocated width and masked by the context afterwards
  const int32_t* bt = p.block_tables + (int64_t)it.seq * p.max_blocks_per_seq + it.blk0;
  const int room = min(part_blocks, p.max_blocks_per_seq - it.blk0);
  const int b0 = lane < room ? bt[lane] : 0;
  const int b1 = 32 + lane < room ? bt[32 + lane] : 0;
  it.ctx = p.context_lens[it.seq];
  const int nblk_ctx = (it.ctx + BS - 1) / BS;
  it.nparts = (it.ctx + part_tokens - 1) / part_tokens;
  const int blk1 = min(it.blk0 + part_blocks, nblk_ctx);
  it.ntiles = max(blk1 - it.blk0, 0);
  it.bt0 = lane < it.ntiles ? b0 : 0;
  it.bt1 = 32 + lane < it.ntiles ? b1 : 0;
}

// ---- the kernel ----------------------------------------------------------------------------------
// Resident CTAs per SM the register allocation aims at (0 = the compiler's default, which lands
// on 168 registers = three 4-warp CTAs for the tensor-core path).  The e5m2 tiles are half the
// size, four CTAs fit in shared memory: VLTR_ATTN_FP8_MINB=4 caps those kernels at 128 registers
// (a few spilled values per tile, L1-resident) for 16 instead of 12 resident warps.
#ifndef VLTR_ATTN_FP8_MINB
#define VLTR_ATTN_FP8_MINB 0
#endif
template <typename T, typename CT, int GT, bool MMA>
constexpr int min_ctas_per_sm() {
  return (MMA && GT == 8 && !std::is_same<T, CT>::value) ? VLTR_ATTN_FP8_MINB : 0;
}

template <typename T, typename CT, int D, int BS, int GT, bool MMA>
__global__ void __launch_bounds__(128, min_ctas_per_sm<T, CT, GT, MMA>()) paged_attention_kernel(const __grid_constant__ Params p) {
  using C = Cfg<T, CT, D, BS, GT, MMA>;
  using QS = typename C::QS;
  using MT = typename C::MT;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int S = p.stages;
#if VLTR_ATTN_TRACE
  long long* tr = p.trace != nullptr ? p.trace + ((int64_t)blockIdx.x * nwarps + warp) * 16 : nullptr;
#else
  constexpr long long* tr = nullptr;   // tracing compiled out: every `if (tr != nullptr)` folds away
#endif
  if (tr != nullptr && lane == 0) tr[0] = gtime();
  int tr_items = 0, tr_tiles = 0;
  long long tr_wait = 0, tr_loop0 = 0, tr_loop = 0;
  const int wbytes = warp_smem_bytes<C>(S);
  unsigned char* wbase = smem_raw + (size_t)warp * wbytes;
  const uint32_t tiles_s = smem_u32(wbase);                        // S x {K tile, V tile}
  QS* q_s = reinterpret_cast<QS*>(wbase + S * 2 * C::TILE_BYTES);  // [GT][D]      (FMA path)
  QS* p_s = reinterpret_cast<QS*>(reinterpret_cast<unsigned char*>(q_s) +
                                  ((C::Q_BYTES + 15) / 16) * 16);  // [GT][BS]     (FMA path)
  const uint32_t bars = smem_u32(reinterpret_cast<unsigned char*>(p_s) +
                                 ((C::P_BYTES + 15) / 16) * 16);   // S mbarriers
  const uint32_t q_sa = smem_u32(q_s), p_sa = smem_u32(p_s);
  int* cum = reinterpret_cast<int*>(smem_raw + (size_t)nwarps * wbytes);   // [kMaxPlanSeqs + 1]
  int* hist = cum + kMaxPlanSeqs + 1;                                       // [4 x 66 + 8]
  unsigned short* tail_order = reinterpret_cast<unsigned short*>(hist + 272);   // [kMaxPlanSeqs]
  Plan plan{cum, tail_order, 0};

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int part_blocks = p.part_blocks;
  int part_tokens = part_blocks * BS;
  int total_items;
  if (p.use_plan) {
    if (p.adaptive) {
      // size the partitions from the ACTUAL contexts (the host only knows the maximum): about
      // 1.5 work items per resident warp, between the host's minimum and cap.  Every CTA
      // computes the same value (integer arithmetic on the same inputs).
      if (threadIdx.x == 0) hist[131] = 0;
      __syncthreads();
      int mine = 0;
      for (int s = threadIdx.x; s < p.num_seqs; s += blockDim.x)
        mine += (p.context_lens[s] + BS - 1) / BS;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
      if (lane == 0 && mine != 0) atomicAdd(&hist[131], mine);
      __syncthreads();
      const long long tiles = (long long)hist[131] * p.num_kv_heads * p.passes;
      // Many groups (more than half a warp-wave of (sequence, kv head, pass) groups): about 1.5
      // items per resident warp, so the longest-first queue can balance.  Few groups (TP shards
      // with one or two kv heads, small batches): everything in ONE wave - every warp streams
      // its few tiles exactly once and nobody queues behind a long item.  Every group adds about
      // half an item of rounding (its partial last partition).  Measured: profiles/
      // r02_attn_split.txt.
      const long long warps_all = (long long)gridDim.x * nwarps;
      const long long groups = (long long)p.num_seqs * p.num_kv_heads * p.passes;
      long long want_items = (warps_all * 3) / 2;
      if (groups * 2 <= warps_all) {
        want_items = warps_all - groups / 2;
        if (want_items < warps_all / 4) want_items = warps_all / 4;
      }
      long long pb = (tiles + want_items - 1) / want_items;
      if (pb < p.part_blocks) pb = p.part_blocks;
      if (pb > p.part_blocks_cap) pb = p.part_blocks_cap;
      part_blocks = (int)pb;
      part_tokens = part_blocks * BS;
      __syncthreads();
    }
    // partition plan (every CTA builds its own copy; num_seqs <= 1024): prefix sums of the FULL
    // partitions per sequence, and the sequences with a partial last partition counting-sorted
    // by its length, longest first, sequence order inside a bucket (every CTA must build the SAME
    // order).  This runs in every CTA of every launch, in front of the first KV request, so it is
    // kept to one L2 round trip (the context lengths) and three barriers: each warp takes a
    // contiguous range of sequences, counts its buckets in its own histogram row with
    // match_any, warp 0 turns the rows into start offsets (bucket descending, then warp) and
    // scans the full-partition counts, then every warp scatters its own sequences.
    constexpr int NB = 66;                                     // buckets 0..65 (tail_bucket)
    for (int i = threadIdx.x; i < 4 * NB + 8; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int nchunks = (p.num_seqs + 31) >> 5;
    const int per_warp = (nchunks + nwarps - 1) / nwarps;      // chunks of 32 sequences per warp
    const int chunk0 = warp * per_warp, chunk1 = min(chunk0 + per_warp, nchunks);
    int* my_hist = hist + (warp & 3) * NB;
    for (int ch = chunk0; ch < chunk1; ++ch) {
      const int s = ch * 32 + lane;
      int bkt = -1;
      if (s < p.num_seqs) {
        const int c = p.context_lens[s];
        cum[s + 1] = c / part_tokens;
        const int rem = c % part_tokens;
        if (rem != 0 || c == 0)         // empty context: a zero-tile tail item writes zeros
          bkt = tail_bucket(c, rem, part_tokens, BS);
      }
      const unsigned same = __match_any_sync(0xffffffffu, bkt);
      if (bkt >= 0 && lane == __ffs(same) - 1) my_hist[bkt] += __popc(same);
      __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
      // start offsets: position q = 65 - bucket ascending; lane owns q = 3*lane .. 3*lane + 2
      int tot[3], before[3][4];
      int mine = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int q = 3 * lane + k;
        tot[k] = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          before[k][w] = tot[k];
          if (q < NB) tot[k] += hist[w * NB + (NB - 1 - q)];
        }
        mine += tot[k];
      }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int start = incl - mine;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int q = 3 * lane + k;
        if (q < NB) {
#pragma unroll
          for (int w = 0; w < 4; ++w) hist[w * NB + (NB - 1 - q)] = start + before[k][w];
        }
        start += tot[k];
      }
      if (lane == 31) hist[4 * NB] = incl;                     // number of tail items
      // inclusive scan of the full-partition counts
      int carry = 0;
      for (int base = 0; base < p.num_seqs; base += 32) {
        const int s = base + lane;
        const int n = s < p.num_seqs ? cum[s + 1] : 0;
        int inc2 = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc2, o);
          if (lane >= o) inc2 += v;
        }
        __syncwarp();
        if (s < p.num_seqs) cum[s + 1] = carry + inc2;
        carry += __shfl_sync(0xffffffffu, inc2, 31);
      }
      if (lane == 0) cum[0] = 0;
    }
    __syncthreads();
    const int n_tail = hist[4 * NB];
    // stable scatter: a warp's sequences in order, equal buckets ranked by lane
    for (int ch = chunk0; ch < chunk1; ++ch) {
      const int s = ch * 32 + lane;
      int bkt = -1;
      if (s < p.num_seqs) {
        const int c = p.context_lens[s];            // L1 hit: read a moment ago
        const int rem = c % part_tokens;
        if (rem != 0 || c == 0) bkt = tail_bucket(c, rem, part_tokens, BS);
      }
      const unsigned same = __match_any_sync(0xffffffffu, bkt);
      const int leader = __ffs(same) - 1;
      const int rank = __popc(same & ((1u << lane) - 1u));
      int start = 0;
      if (bkt >= 0 && lane == leader) {
        start = my_hist[bkt];
        my_hist[bkt] = start + __popc(same);
      }
      start = __shfl_sync(0xffffffffu, start, leader);
      if (bkt >= 0) tail_order[start + rank] = (unsigned short)s;
      __syncwarp();
    }
    __syncthreads();
    plan.cum = cum;
    plan.tail_order = tail_order;
    plan.n_full = cum[p.num_seqs];
    total_items = (plan.n_full + n_tail) * p.passes * p.num_kv_heads;
  } else {
    __syncwarp();
    total_items = p.num_seqs * p.n_part * p.passes * p.num_kv_heads;
  }

  // lane geometry (FMA path)
  const int k_tok = lane % BS, k_grp = lane / BS;       // q.k: token, chunk phase
  const int v_row = lane / (C::LPR > 0 ? C::LPR : 1), v_sub = lane % (C::LPR > 0 ? C::LPR : 1);
  // lane geometry (MMA path): g = row of the 16 x 8 fragment, t4 = column pair
  const int g = lane >> 2, t4 = lane & 3;
  // e5m2 tensor-core path, block size >= 16: row r of a 16-token M tile of S^T = K.Q^T holds token
  //   r = 2t + e      (r < 8)  -> 4t + e          r = 8 + 2t + e  -> 4t + 2 + e      (t = 0..3, e = 0,1)
  // so that after the movmatrix transpose the k slots {2t4, 2t4+1, 2t4+8, 2t4+9} of P^T are the
  // CONSECUTIVE tokens 4t4 .. 4t4+3 - four adjacent bytes of a V row.  prow0 = token of row g,
  // prow0 + prow8 = token of row g + 8.  Everything else keeps the identity (prow0 = g, prow8 = 8).
  constexpr bool PERM = MMA && C::FP8 && BS >= 16;
  const int prow0 = PERM ? 4 * (g >> 1) + (g & 1) : g;
  constexpr int prow8 = PERM ? 2 : 8;

  int stage_c = 0;        // consumer cursor over the ring (continues across items)
  uint32_t phase_c = 0;
  int stage_p = 0;        // producer cursor
  int inflight = 0;       // tiles issued and not yet consumed

  const float scale_l2 = p.scale * kLog2e;
  const int total_warps = gridDim.x * nwarps;

  // issue tiles of `it` while the ring has room
  auto pump = [&](Item& it) {
    if (!it.valid) return;
    const char* kbase = reinterpret_cast<const char*>(p.k_cache) + (int64_t)it.kvh * p.kv_head_stride;
    const char* vbase = reinterpret_cast<const char*>(p.v_cache) + (int64_t)it.kvh * p.kv_head_stride;
    while (inflight < S && it.issued < it.ntiles) {
      const int i = it.issued;
      const int bid = __shfl_sync(0xffffffffu, i < 32 ? it.bt0 : it.bt1, i & 31);
      if (lane == 0) {
        const uint32_t bar = bars + 8 * stage_p;
        const uint32_t dst = tiles_s + stage_p * 2 * C::TILE_BYTES;
        mbar_expect_tx(bar, 2 * C::TILE_BYTES);
        bulk_g2s(dst, kbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES, bar);
        bulk_g2s(dst + C::TILE_BYTES, vbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES, bar);
      }
      stage_p = stage_p + 1 == S ? 0 : stage_p + 1;
      ++it.issued;
      ++inflight;
    }
  };

  // Work-item lookahead.  The queue ticket for the NEXT item is drawn (async atomic) when the
  // current item has <= S+2 tiles left, and turned into an item (context length + block ids) once
  // all tiles of the current item are in flight - late enough not to hoard work other warps could
  // take, early enough that its tiles are landing while the current item finishes.
  // Everything above reads only step inputs (context lengths) and builds shared-memory state:
  // with programmatic dependent launch it ran while the previous kernel (reshape_and_cache, which
  // writes the newest token's K/V) was still draining.  After the wait the query and the KV cache
  // are read.
  // The first work item is claimed here too, still ahead of the wait: the queue counter of this
  // launch was zeroed by the launch that used the slot kCounterSlots launches ago, and an item is
  // made of step inputs only (context length, block ids) - two L2 round trips that would
  // otherwise sit between the end of the previous kernel and the first KV request.
  Item cur, nxt;
  nxt.valid = 0;
  int pend = blockIdx.x * nwarps + warp;      // lane 0's value is the one claim() broadcasts
  bool pend_valid = false, nxt_ready = false;
  claim(cur, pend, p, plan, total_items, part_blocks, BS, lane);
  pdl_wait();
  pdl_trigger();
  if (tr != nullptr && lane == 0) tr[1] = gtime();
  if (p.plan_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.plan_out[0] = part_blocks;
  pump(cur);
  auto advance = [&]() {
    if (!nxt_ready) {
      if (!pend_valid) pend = fetch_id(p, lane, total_warps);
      claim(nxt, pend, p, plan, total_items, part_blocks, BS, lane);
    }
    cur = nxt;
    nxt.valid = 0;
    nxt_ready = false;
    pend_valid = false;
    pump(cur);
  };

  while (cur.valid) {
    const int seq = cur.seq, part = cur.part, kvh = cur.kvh, ctx = cur.ctx;
    const int ntiles = cur.ntiles, blk0 = cur.blk0;
    const int nblk_ctx = (ctx + BS - 1) / BS;
    const int head0 = kvh * p.group + cur.pass * GT;             // first q head of this pass
    const int nheads = min(GT, p.group - cur.pass * GT);          // live heads in this pass
    const bool partial = p.always_partial || cur.nparts > 1;

    if (ntiles <= 0) {
      // empty context: the reference leaves zeros (its accumulators stay 0)
      if (part == 0) {
        T* o = reinterpret_cast<T*>(p.out) + ((int64_t)seq * p.num_heads + head0) * D;
        for (int i = lane; i < nheads * D; i += 32) o[i] = ST<T>::from_f(0.f);
      }
      advance();
      continue;
    }

    // running softmax state.  FMA: per head, replicated over the lanes of a token / row.
    // MMA: both products are computed TRANSPOSED, so that the 16-row M dimension of
    // mma.m16n8k16 is always full and the 8-wide N dimension carries the query heads:
    //   S^T[token, head] = K[token, :] . Q^T[:, head]     A = K tile rows,  B = q fragments
    //   O^T[d, head]    += V^T[d, token] . P^T[token, head]   A = V tile rows,  B = P^T
    // (the head-major orientation pads 8 heads to 16 rows: twice the mma count, zero operand
    // registers to shuffle, and 64 instead of 32 accumulator registers for O).  A thread owns
    // head columns 8nh + 2t4 + {0,1} (replicated over the 8 row groups g): HC of them.
    constexpr int NH = MMA ? GT / 8 : 1;              // 8-head N tiles
    constexpr int HC = 2 * NH;                        // head columns per thread
    constexpr int MTL = MMA ? (BS + 15) / 16 : 1;     // 16-token M tiles per KV block
    constexpr int DT = D / 16;                        // 16-row tiles of the head dimension
    constexpr int NR = MMA ? HC : GT;
    float m[NR], l[NR], slope[NR];
    float acc[MMA ? 1 : GT][MMA ? 1 : C::RI];
    float o_acc[MMA ? DT * NH : 1][4];   // [dt][nh]: (d = 16dt+g, heads 2t4,2t4+1), (d + 8, same heads)
    uint32_t qa[MMA ? C::KS : 1][4];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      m[r] = -INFINITY;
      l[r] = 0.f;
      const int h = MMA ? 8 * (r >> 1) + 2 * t4 + (r & 1) : r;
      slope[r] = (p.alibi_slopes != nullptr && h < nheads) ? p.alibi_slopes[head0 + h] * kLog2e : 0.f;
    }
    if constexpr (MMA) {
#pragma unroll
      for (int i = 0; i < DT * NH; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
      // q fragments: B operand of S^T = K . Q^T; for head tile nh the pair (qa[ks][nh],
      // qa[ks][2 + nh]) holds Q[head 8nh + g][k slots 2t4,2t4+1 | 2t4+8,2t4+9] of k-step ks.  For
      // e5m2 caches the k slots of a step are permuted so that a thread's four K bytes are
      // contiguous: slots {2t,2t+1,2t+8,2t+9} <-> head-dim offsets {4t,4t+1,4t+2,4t+3}.
      const T* qg = reinterpret_cast<const T*>(p.q) + (int64_t)seq * p.q_stride + (int64_t)head0 * D;
#pragma unroll
      for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int h = g + 8 * r;
          uint32_t lo = 0, hi = 0;
          if (h < nheads && (r == 0 || C::R16)) {
            const T* src = qg + (int64_t)h * D + ks * 16;
            if constexpr (C::FP8) {
              const uint2 w = *reinterpret_cast<const uint2*>(src + 4 * t4);
              if constexpr (std::is_same<T, __half>::value) {
                lo = w.x;
                hi = w.y;
              } else {   // bf16 query, f16 tensor-core operands
                const float2 a = unpack2<__nv_bfloat16>(w.x), b = unpack2<__nv_bfloat16>(w.y);
                lo = pack2<__half>(a.x, a.y);
                hi = pack2<__half>(b.x, b.y);
              }
            } else {
              lo = *reinterpret_cast<const uint32_t*>(src + 2 * t4);
              hi = *reinterpret_cast<const uint32_t*>(src + 8 + 2 * t4);
            }
          }
          qa[ks][r] = lo;        // b0 of head tile r: k slots 2t, 2t+1
          qa[ks][2 + r] = hi;    // b1 of head tile r: k slots 2t+8, 2t+9
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < GT; ++h)
#pragma unroll
        for (int i = 0; i < C::RI; ++i) acc[h][i] = 0.f;
      // stage q: [GT][D] as QS (zero rows for dead heads so their dot products stay finite)
      const T* qg = reinterpret_cast<const T*>(p.q) + (int64_t)seq * p.q_stride + (int64_t)head0 * D;
      for (int i = lane; i < GT * D; i += 32) {
        const int h = i / D;
        const float f = h < nheads ? ST<T>::to_f(qg[i]) : 0.f;
        q_s[i] = to_qs<QS>(f);
      }
      __syncwarp();
    }

    if (tr != nullptr) {
      tr_loop0 = clock64();
      if (lane == 0) tr[8] = gtime();       // item set up (claimed, q staged)
    }
    // Tensor-core path, e5m2 cache: TWO KV blocks per loop iteration when the item has them.
    // One warp's tile is a dependent chain (shared loads -> byte permutes -> mma chain -> shuffle
    // max -> exp -> transpose -> mma) of ~1800 cycles whatever its instruction count
    // (profiles/r02_attn_timeline.txt); two blocks give the warp two independent chains to
    // interleave and one softmax update (shuffles, rescale test) per 32 tokens.  NTP is the
    // compile-time capacity, `nt` (warp-uniform) what this iteration really has; the second
    // block's code is skipped (QK, PV) or masked (softmax rows) when nt == 1.
    constexpr int NTP =
        (MMA && MTL == 1 && BS >= 16 && VLTR_ATTN_PAIR >= (C::FP8 ? 1 : 2)) ? 2 : 1;
    constexpr int MT2 = NTP * MTL;
    for (int tile = 0; tile < ntiles;) {
      const int nt = (NTP == 2 && ntiles - tile >= 2) ? 2 : 1;
      if (!pend_valid && !nxt_ready && ntiles - tile <= S + 2) {
        pend = fetch_id(p, lane, total_warps);
        pend_valid = true;
      } else if (pend_valid && !nxt_ready && cur.issued >= ntiles) {
        claim(nxt, pend, p, plan, total_items, part_blocks, BS, lane);
        nxt_ready = true;
        pend_valid = false;
        pump(nxt);
      }
      // ring stages of the iteration's blocks (consecutive, wrapping)
      const int stage_1 = stage_c + 1 == S ? 0 : stage_c + 1;
      const uint32_t phase_1 = stage_c + 1 == S ? phase_c ^ 1u : phase_c;
      const uint32_t ktile = tiles_s + stage_c * 2 * C::TILE_BYTES;
      const uint32_t vtile = ktile + C::TILE_BYTES;
      const uint32_t ktile1 = tiles_s + stage_1 * 2 * C::TILE_BYTES;   // used when NTP == 2
      const uint32_t kt[2] = {ktile, ktile1};
      const uint32_t vt_[2] = {vtile, ktile1 + (uint32_t)C::TILE_BYTES};
      long long tw0 = 0;
      if (tr != nullptr) tw0 = clock64();
      mbar_wait(bars + 8 * stage_c, phase_c);
      if (NTP == 2 && nt == 2) mbar_wait(bars + 8 * stage_1, phase_1);
      if (tr != nullptr) {
        tr_wait += clock64() - tw0;
        if (tr_tiles == 0 && lane == 0) tr[2] = gtime();
        tr_tiles += nt;
      }
      const int tok0 = (blk0 + tile) * BS;
      // only the context's last block can be partial; it is the iteration's last block
      const int last_u = nt - 1;
      const bool last_partial = (blk0 + tile + last_u == nblk_ctx - 1) && (ctx % BS != 0);

      if constexpr (MMA) {
        // ---- tensor-core path -------------------------------------------------------------------
        if (last_partial) {
          // zero the out-of-context columns of the V tile (they may hold NaNs, p = 0 there)
          const int valid = ctx - (tok0 + last_u * BS);
          const int stage_l = last_u == 0 ? stage_c : stage_1;
          unsigned char* vt = wbase + (size_t)stage_l * 2 * C::TILE_BYTES + C::TILE_BYTES;
          for (int i = lane; i < D * BS; i += 32)
            if (i % BS >= valid) reinterpret_cast<CT*>(vt)[i] = CT(0);
          // these generic-proxy writes must be ordered before the TMA refill of this stage
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
        }
        const int j = lane >> 3;      // ldmatrix: lane supplies row (lane & 7) of matrix j
        // S^T: (token g | g+8 of M tile mt, heads 2t4,2t4+1 of tile nh); M tile mt belongs to
        // block u = mt / MTL of the iteration
        float st[MT2][NH][4];
#pragma unroll
        for (int mt = 0; mt < MT2; ++mt)
#pragma unroll
          for (int nh = 0; nh < NH; ++nh) st[mt][nh][0] = st[mt][nh][1] = st[mt][nh][2] = st[mt][nh][3] = 0.f;
        // the KS products of a tile accumulate into the same registers: a dependent chain of KS
        // tensor-core latencies per warp.  SPLIT_QK keeps a second accumulator for the odd
        // k-steps (two chains of KS / 2) and adds them at the end (not needed with two blocks
        // per iteration: they are two chains already).
        constexpr bool SPLIT_QK = VLTR_ATTN_SPLIT_QK && NTP == 1 && C::KS >= 4 && C::KS % 2 == 0;
        float st2[SPLIT_QK ? MTL : 1][SPLIT_QK ? NH : 1][4];
        if constexpr (SPLIT_QK) {
#pragma unroll
          for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
            for (int nh = 0; nh < NH; ++nh)
              st2[mt][nh][0] = st2[mt][nh][1] = st2[mt][nh][2] = st2[mt][nh][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
          for (int mt = 0; mt < MT2; ++mt) {
            const int u = mt / MTL, ms = mt % MTL;      // block of the iteration, M tile inside it
            if (NTP == 2 && u >= nt) continue;          // warp-uniform
            const uint32_t kbase_ = kt[u];
            uint32_t a[4];
            if constexpr (!C::FP8) {
              // K tile [D/8][BS][8]: (chunk c, tokens 8b..8b+7) is one 8x8 b16 matrix, rows 16 B
              // apart.  A fragment: a0 = (tokens 0-7, k 0-7), a1 = (tokens 8-15, k 0-7),
              // a2 = (tokens 0-7, k 8-15), a3 = (tokens 8-15, k 8-15)
              if constexpr (BS >= 16) {
                ldsm_x4(kbase_ + (((2 * ks + (j >> 1)) * BS + 16 * ms + 8 * (j & 1) + (lane & 7)) << 4), a);
              } else {   // BS == 8: tokens 8..15 of the M tile do not exist
                ldsm_x4(kbase_ + (((2 * ks + (j >> 1)) * BS + (lane & 7)) << 4), a);
                a[1] = 0u;
                a[3] = 0u;
              }
            } else {
              // e5m2 K tile [D/16][BS][16]: bytes 4t4..4t4+3 of (k-step ks, token) are this
              // thread's four k slots.  Which token sits in which row of the M tile is free
              // (the softmax runs down the rows, P^T inherits the order): rows g and g + 8 take
              // tokens prow0 and prow0 + 2 (see PERM above) so that V needs one load per row.
              const uint32_t w0 = lds4(kbase_ + ((ks * BS + 16 * ms + prow0) << 4) + 4 * t4);
              a[0] = __byte_perm(w0, 0, 0x1404);
              a[2] = __byte_perm(w0, 0, 0x3424);
              if constexpr (BS >= 16) {
                const uint32_t w1 = lds4(kbase_ + ((ks * BS + 16 * ms + prow0 + prow8) << 4) + 4 * t4);
                a[1] = __byte_perm(w1, 0, 0x1404);
                a[3] = __byte_perm(w1, 0, 0x3424);
              } else {
                a[1] = 0u;
                a[3] = 0u;
              }
            }
#pragma unroll
            for (int nh = 0; nh < NH; ++nh) {
              if (SPLIT_QK && (ks & 1))
                mma16816<MT>(st2[SPLIT_QK ? mt : 0][SPLIT_QK ? nh : 0], a, qa[ks][nh], qa[ks][2 + nh]);
              else
                mma16816<MT>(st[mt][nh], a, qa[ks][nh], qa[ks][2 + nh]);
            }
          }
        }
        if constexpr (SPLIT_QK) {
#pragma unroll
          for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
            for (int nh = 0; nh < NH; ++nh)
#pragma unroll
              for (int i = 0; i < 4; ++i) st[mt][nh][i] += st2[mt][nh][i];
        }
        // ---- online softmax down the token rows of each head column (log2 domain) --------------
        // only the context's last block can hold out-of-context tokens, and ALiBi is rare: the
        // common tile takes the short form (bit-identical: slope == 0 adds +0)
        const bool plain = !last_partial && p.alibi_slopes == nullptr && BS >= 16 && nt == NTP;
        uint32_t pb[MT2][NH][2];      // P^T as B operand: [0] tokens 2t4,2t4+1; [1] tokens 8+2t4,9+2t4
#pragma unroll
        for (int nh = 0; nh < NH; ++nh) {
          float m_new[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * nh + e;
            float mx = -INFINITY;
#pragma unroll
            for (int mt = 0; mt < MT2; ++mt) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float sv = st[mt][nh][2 * hh + e] * scale_l2;
                if (!plain) {
                  const int u = mt / MTL;
                  const int row = 16 * (mt % MTL) + prow8 * hh + prow0;
                  const int tok = tok0 + u * BS + row;
                  sv += slope[c] * (float)(tok - ctx + 1);
                  sv = (u < nt && tok < ctx && row < BS) ? sv : -INFINITY;
                }
                st[mt][nh][2 * hh + e] = sv;
                mx = fmaxf(mx, sv);
              }
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
            m_new[e] = fmaxf(m[c], mx);     // finite: a scheduled tile has a valid token
            // the running maximum settles after the first tiles: rescale only when some head's
            // maximum moved (warp-uniform; alpha == 1 exactly otherwise)
            if (__any_sync(0xffffffffu, m_new[e] != m[c])) {
              const float alpha = ex2_ftz(m[c] - m_new[e]);
              m[c] = m_new[e];
              l[c] *= alpha;
#pragma unroll
              for (int dt = 0; dt < DT; ++dt) {
                o_acc[dt * NH + nh][e] *= alpha;
                o_acc[dt * NH + nh][2 + e] *= alpha;
              }
            }
          }
#pragma unroll
          for (int mt = 0; mt < MT2; ++mt) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (BS >= 16 || hh == 0) {
                // (token row, heads 2t4 and 2t4+1) packed, rounded to the operand type; the row
                // sums use the rounded weights, like the products do
                const uint32_t pk = pack2<MT>(ex2_ftz(st[mt][nh][2 * hh] - m_new[0]),
                                              ex2_ftz(st[mt][nh][2 * hh + 1] - m_new[1]));
                const float2 pf = unpack2<MT>(pk);
                l[2 * nh] += pf.x;
                l[2 * nh + 1] += pf.y;
                // 8x8 transpose across the warp: (token g, heads 2t4..) -> (head g, tokens 2t4..)
                uint32_t tr_;
                asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(tr_) : "r"(pk));
                pb[mt][nh][hh] = tr_;
              } else {
                pb[mt][nh][hh] = 0u;
              }
            }
          }
        }
        // ---- O^T += V^T . P^T -------------------------------------------------------------------
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
#pragma unroll
          for (int mt = 0; mt < MT2; ++mt) {
            const int u = mt / MTL, ms = mt % MTL;
            if (NTP == 2 && u >= nt) continue;          // warp-uniform: stale stage, p = 0
            const uint32_t vbase_ = vt_[u];
            uint32_t a[4];
            if constexpr (!C::FP8) {
              // V tile [D][BS]: (rows 8b..8b+7, tokens 8c..8c+7) is one 8x8 matrix, rows 2*BS B
              // apart.  a0 = (d 0-7, tok 0-7), a1 = (d 8-15, tok 0-7), a2 = (d 0-7, tok 8-15), a3
              if constexpr (BS >= 16) {
                ldsm_x4(vbase_ + ((16 * dt + 8 * (j & 1) + (lane & 7)) * BS + 16 * ms + 8 * (j >> 1)) * 2, a);
              } else {
                ldsm_x4(vbase_ + ((16 * dt + 8 * (j & 1) + (lane & 7)) * BS) * 2, a);
                a[2] = 0u;
                a[3] = 0u;
              }
            } else {
              if constexpr (PERM) {
                // e5m2 V tile [D][BS] bytes.  With the token rows of S^T permuted, the P^T
                // fragments carry tokens 4t4, 4t4+1 (k slots 2t4, 2t4+1) and 4t4+2, 4t4+3 (k slots
                // 2t4+8, 2t4+9): the four V bytes of a row are ONE 32-bit load
                const uint32_t row0 = vbase_ + (16 * dt + g) * BS + 16 * ms + 4 * t4;
                const uint32_t w0 = lds4(row0), w1 = lds4(row0 + 8 * BS);
                a[0] = __byte_perm(w0, 0, 0x1404);
                a[1] = __byte_perm(w1, 0, 0x1404);
                a[2] = __byte_perm(w0, 0, 0x3424);
                a[3] = __byte_perm(w1, 0, 0x3424);
              } else {
                // block size 8: two 16-bit loads (tokens 2t4, 2t4+1 of rows g, g + 8)
                const uint32_t row0 = vbase_ + (16 * dt + g) * BS + 16 * ms + 2 * t4;
                uint32_t w0, w1;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w0) : "r"(row0));
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w1) : "r"(row0 + 8 * BS));
                a[0] = __byte_perm(w0, 0, 0x1404);
                a[1] = __byte_perm(w1, 0, 0x1404);
                a[2] = 0u;
                a[3] = 0u;
              }
            }
#pragma unroll
            for (int nh = 0; nh < NH; ++nh)
              mma16816<MT>(o_acc[dt * NH + nh], a, pb[mt][nh][0], pb[mt][nh][1]);
          }
        }
      } else {
        // ---- FMA path: s = q.k for this lane's token ---------------------------------------------
        float s[GT];
#pragma unroll
        for (int h = 0; h < GT; ++h) s[h] = 0.f;
#pragma unroll
        for (int j = 0; j < C::KJ; ++j) {
          const int c = k_grp + j * C::TG;
          if (C::NCH % C::TG == 0 || c < C::NCH) {
            const uint4 kv = lds16(ktile + (c * BS + k_tok) * 16);
#pragma unroll
            for (int h = 0; h < GT; ++h) {
              uint4 qw[C::QW];
#pragma unroll
              for (int w = 0; w < C::QW; ++w)
                qw[w] = lds16(q_sa + (h * D + c * C::EPC) * (int)sizeof(QS) + w * 16);
              s[h] = dot_chunk<T, C>(s[h], kv, qw);
            }
          }
        }
#pragma unroll
        for (int h = 0; h < GT; ++h) {
#pragma unroll
          for (int o = BS; o < 32; o <<= 1) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
        }
        // online softmax (log2 domain)
        const int tok = tok0 + k_tok;
        const bool valid = tok < ctx;
#pragma unroll
        for (int h = 0; h < GT; ++h) {
          float sv = s[h] * scale_l2 + slope[h] * (float)(tok - ctx + 1);
          sv = valid ? sv : -INFINITY;
          float mx = sv;
#pragma unroll
          for (int o = 1; o < BS; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          const float m_new = fmaxf(m[h], mx);
          const QS pq = to_qs<QS>(exp2f(sv - m_new));
          if (k_grp == 0) p_s[h * BS + k_tok] = pq;
          if (m_new > m[h]) {  // warp-uniform
            const float alpha = exp2f(m[h] - m_new);
            l[h] *= alpha;
#pragma unroll
            for (int i = 0; i < C::RI; ++i) acc[h][i] *= alpha;
            m[h] = m_new;
          }
          l[h] += qs_to_f<QS>(pq);
        }
        __syncwarp();
        // acc += p.v
        const int vvalid = ctx - tok0 - v_sub * C::EPC;  // valid tokens in my chunk
        uint4 pw[C::P_IN_REGS ? GT : 1][C::QW];
        if constexpr (C::P_IN_REGS) {
#pragma unroll
          for (int h = 0; h < GT; ++h)
#pragma unroll
            for (int w = 0; w < C::QW; ++w)
              pw[h][w] = lds16(p_sa + (h * BS + v_sub * C::EPC) * (int)sizeof(QS) + w * 16);
        }
#pragma unroll
        for (int i = 0; i < C::RI; ++i) {
          const int row = v_row + i * C::RPI;
          if (D % C::RPI == 0 || row < D) {
            uint4 vv = lds16(vtile + (row * C::LPR + v_sub) * 16);
            if (last_partial) mask_chunk<C>(vv, vvalid);
#pragma unroll
            for (int h = 0; h < GT; ++h) {
              if constexpr (C::P_IN_REGS) {
                acc[h][i] = dot_chunk<T, C>(acc[h][i], vv, pw[h]);
              } else {
                uint4 pl[C::QW];
#pragma unroll
                for (int w = 0; w < C::QW; ++w)
                  pl[w] = lds16(p_sa + (h * BS + v_sub * C::EPC) * (int)sizeof(QS) + w * 16);
                acc[h][i] = dot_chunk<T, C>(acc[h][i], vv, pl);
              }
            }
          }
        }
      }
      __syncwarp();  // every lane is done with this iteration's stages (and with p_s)
      for (int u = 0; u < nt; ++u) {
        --inflight;
        if (++stage_c == S) {
          stage_c = 0;
          phase_c ^= 1;
        }
      }
      tile += nt;
      // refill the freed stages: this item first, then the next one (lookahead across items)
      pump(cur);
      if (nxt_ready) pump(nxt);
    }

    if (tr != nullptr) {
      tr_loop += clock64() - tr_loop0;
      if (lane == 0) tr[9] = gtime();       // tiles done
    }
    // ---- epilogue: finish the reductions, normalise, store / publish the partial ----------------
    // Layout-independent description of what this lane owns: FMA: head h, rows v_row + i*RPI
    // (lanes v_sub == 0); MMA: rows g (+8), columns 8nd + 2t4 + {0,1}.
    if constexpr (MMA) {
      // every row group g holds partial row sums of its own tokens
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 4);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 8);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 16);
      }
    } else {
#pragma unroll
      for (int h = 0; h < GT; ++h) {
        float lt = l[h];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) lt += __shfl_xor_sync(0xffffffffu, lt, o);
        l[h] = lt * (1.0f / C::TG);  // every token was counted by its TG lanes
#pragma unroll
        for (int i = 0; i < C::RI; ++i) {
          float a = acc[h][i];
#pragma unroll
          for (int o = 1; o < C::LPR; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
          acc[h][i] = a;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int h = MMA ? 8 * (r >> 1) + 2 * t4 + (r & 1) : r;
      if (h < nheads) {
        const int head = head0 + h;
        const float inv = __fdividef(1.f, l[r] + 1e-6f);
        const int64_t pi = ((int64_t)seq * p.num_heads + head) * p.n_part + part;
        if (partial && (MMA ? g == 0 : lane == 0)) {
          p.max_logits[pi] = m[r] * kLn2;  // natural-log domain, as the reference stores qk_max
          p.exp_sums[pi] = l[r];
        }
        T* o = reinterpret_cast<T*>(p.out) + ((int64_t)seq * p.num_heads + head) * D;
        const bool direct = cur.nparts == 1;   // single partition: the result is final
        if constexpr (MMA) {
          // this thread's column of O^T: rows d = 16dt + g and 16dt + g + 8 (the 8 row groups of
          // a head column write 8 consecutive elements)
#pragma unroll
          for (int dt = 0; dt < DT; ++dt) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int d = 16 * dt + 8 * hh + g;
              const float v = o_acc[dt * NH + (r >> 1)][2 * hh + (r & 1)] * inv;
              if (direct) o[d] = ST<T>::from_f(v);
              if (partial) {
                if (p.tmp_is_f32)
                  reinterpret_cast<float*>(p.tmp_out)[pi * D + d] = v;
                else
                  reinterpret_cast<T*>(p.tmp_out)[pi * D + d] = ST<T>::from_f(v);
              }
            }
          }
        } else if (v_sub == 0) {
#pragma unroll
          for (int i = 0; i < C::RI; ++i) {
            const int row = v_row + i * C::RPI;
            if (D % C::RPI == 0 || row < D) {
              const float v = acc[r][i] * inv;
              if (direct) o[row] = ST<T>::from_f(v);
              if (partial) {
                if (p.tmp_is_f32)
                  reinterpret_cast<float*>(p.tmp_out)[pi * D + row] = v;
                else
                  reinterpret_cast<T*>(p.tmp_out)[pi * D + row] = ST<T>::from_f(v);
              }
            }
          }
        }
      }
    }

    if (tr != nullptr && lane == 0) tr[10] = gtime();   // result / partial stored
    // ---- merge: the warp that publishes a sequence's LAST partition combines them all ------------
    // (reference: paged_attention_v2_reduce_kernel, :511-602: weights exp_sum_j*exp(max_j-max) /
    // (sum + 1e-6) over per-partition normalised outputs)
    if (cur.nparts > 1 && !p.defer_merge) {
      __threadfence();
      __syncwarp();
      const int gid = (seq * p.passes + cur.pass) * p.num_kv_heads + kvh;
      int done = 0;
      if (lane == 0) done = atomicAdd(&p.tickets[gid], 1);
      done = __shfl_sync(0xffffffffu, done, 0);
      if (done == cur.nparts - 1) {
        __threadfence();
        merge_partials<T, D>(p, seq, head0, nheads, cur.nparts, lane);
        if (lane == 0) p.tickets[gid] = 0;   // leave the ticket table clean for the next launch
      }
    }
    __syncwarp();  // q_s / p_s are rewritten by the next item
    ++tr_items;
    if (tr != nullptr && lane == 0) {
      tr[11] = gtime();                     // merge (if any) done
      tr[12] = cur.nparts;
      tr[13] = ntiles;
    }

    advance();
  }
  if (tr != nullptr && lane == 0) {
    tr[3] = gtime();
    tr[4] = tr_items;
    tr[5] = tr_tiles;
    tr[6] = tr_wait;      // SM cycles spent waiting for tiles
    tr[7] = tr_loop;      // SM cycles in the tile loops (waiting + computing)
  }

  // self-resetting queue: the last warp to leave zeroes the counters for the next launch
  if (lane == 0) {
    const int done = atomicAdd(&p.counters[1], 1);
    if (done == total_warps - 1) {
      p.counters[0] = 0;
      p.counters[1] = 0;
      __threadfence();
    }
  }
}

// ---- second launch of the latency-bound regime ---------------------------------------------------
// When a launch has far fewer (sequence, kv-head) groups than the machine has resident warps
// (tensor-parallel shards: 70B at TP8 is 1 kv head per GPU; small batches), contexts are cut into
// many short partitions so that every warp streams a few KV tiles - and then merging a sequence's
// partitions inside the main kernel would put nheads x nparts dependent L2 round trips on ONE warp
// at the very end of the launch (measured: up to 20 us of a 46 us launch, profiles/
// r01_attn_tune_timeline.txt).  Here the merge is its own grid instead: one warp per
// (sequence, query head), all of them in parallel, launched with programmatic dependent launch so
// that it is resident when the main kernel drains.  Same arithmetic as merge_partials (reference:
// paged_attention_v2_reduce_kernel, attention_kernels.cu:511-602).
template <typename T, int D>
__global__ void __launch_bounds__(128) merge_partials_kernel(const __grid_constant__ Params p, int BS) {
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 4 + (threadIdx.x >> 5);
  pdl_wait();          // the main kernel has completed: partials and plan_out are visible
  pdl_trigger();
  if (idx >= p.num_seqs * p.num_heads) return;
  const int seq = idx / p.num_heads;
  const int part_tokens = __ldcg(p.plan_out) * BS;
  const int ctx = p.context_lens[seq];
  const int np = (ctx + part_tokens - 1) / part_tokens;
  if (np <= 1) return;                       // single partition: the main kernel stored the result
  const int64_t base = (int64_t)idx * p.n_part;
  constexpr int KW = 4;                      // partitions per lane: np <= 128 (host guarantees)
  float mj[KW], sj[KW];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    const int j = lane + 32 * k;
    mj[k] = j < np ? __ldcg(p.max_logits + base + j) : -INFINITY;
    sj[k] = j < np ? __ldcg(p.exp_sums + base + j) : 0.f;
    m = fmaxf(m, mj[k]);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    sj[k] = lane + 32 * k < np ? sj[k] * expf(mj[k] - m) : 0.f;
    sum += sj[k];
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = __fdividef(1.0f, sum + 1e-6f);
  constexpr int NV = (D + 127) / 128;
  float4 a[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int JB = 8;                      // partial rows in flight per lane
  for (int j0 = 0; j0 < np; j0 += JB) {
    float4 v[JB][NV];
    float w[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      const int j = min(j0 + jj, np - 1);
      float wk = 0.f;
#pragma unroll
      for (int k = 0; k < KW; ++k)            // owner lane (j & 31) holds it in slot j >> 5
        if ((j >> 5) == k) wk = __shfl_sync(0xffffffffu, sj[k], j & 31);
      w[jj] = j0 + jj < np ? wk * inv : 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int d = 4 * lane + 128 * i;
        v[jj][i] = (D % 128 == 0 || d < D)
                       ? load_cg_4<T>(p.tmp_out, (base + j) * D + d, p.tmp_is_f32 != 0)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int jj = 0; jj < JB; ++jj)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        a[i].x += v[jj][i].x * w[jj];
        a[i].y += v[jj][i].y * w[jj];
        a[i].z += v[jj][i].z * w[jj];
        a[i].w += v[jj][i].w * w[jj];
      }
  }
  T* o = reinterpret_cast<T*>(p.out) + (int64_t)idx * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d = 4 * lane + 128 * i;
    if (D % 128 == 0 || d < D) {
      o[d] = ST<T>::from_f(a[i].x);
      o[d + 1] = ST<T>::from_f(a[i].y);
      o[d + 2] = ST<T>::from_f(a[i].z);
      o[d + 3] = ST<T>::from_f(a[i].w);
    }
  }
}

}  // namespace attn
}  // namespace vltr
