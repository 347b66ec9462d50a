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
//                                               per-warp mbarrier ring, 128-bit LDS from there
//   logits for the whole context in smem        online softmax per 1 KV block in registers:
//   (4*ctx bytes -> 8K..57K token limit),       no context-length limit, no CTA-wide barrier
//   3 __syncthreads + smem tree per CTA         in the main loop (warps are independent)
//   q.k with cvt f16->f32 + FFMA; p.v in        FHFMA (fma.rn.f32.f16 / .bf16, sm_100 mixed
//   half precision (mul.f16x2)                  precision FMA): f16/bf16 operands straight from
//                                               smem, fp32 accumulation, no conversions
//   static grid, tail effects on ragged         persistent grid (CTAs/SM x 148), warps pull
//   context lengths                             items from an atomic queue (self-resetting)
//
// Bound: HBM.  Algorithmic bytes per call = sum_i L_i*2*Hkv*D*s_cache + q + out + tables.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace vltr {
namespace attn {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct Params {
  void* out;                    // [num_seqs, num_heads, D]  T
  const void* q;                // [num_seqs, num_heads, D]  T, row stride q_stride
  const void* k_cache;
  const void* v_cache;
  const int32_t* block_tables;  // [num_seqs, max_blocks_per_seq]
  const int32_t* context_lens;  // [num_seqs]
  const float* alibi_slopes;    // [num_heads] or null
  // split mode (n_part > 1 or v2): partial results
  float* exp_sums;              // [num_seqs, num_heads, n_part]
  float* max_logits;            // [num_seqs, num_heads, n_part]
  void* tmp_out;                // [num_seqs, num_heads, n_part, D]  T or float
  int split;                    // 0: write `out`, 1: write partials
  int tmp_is_f32;
  int num_seqs, num_heads, num_kv_heads, group, passes;
  int max_blocks_per_seq;
  int n_part, part_blocks;
  float scale;
  int64_t q_stride;             // elements
  int64_t kv_block_stride;      // bytes
  int64_t kv_head_stride;       // bytes
  int total_items;
  int stages;
  int* counters;                // [0] next item, [1] finished warps
};

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

// ---- compile-time geometry ------------------------------------------------------------------------
template <typename T, typename CT, int D_, int BS_, int GT_>
struct Cfg {
  static constexpr int D = D_, BS = BS_, GT = GT_;
  static constexpr bool FP8 = !std::is_same<T, CT>::value;
  static constexpr int EPC = 16 / sizeof(CT);       // elements per 16-byte cache chunk (== x)
  static constexpr int NCH = D / EPC;               // K chunks per token
  static constexpr int TG = 32 / BS;                // lanes sharing one token in q.k
  static constexpr int KJ = (NCH + TG - 1) / TG;    // K chunks per lane
  static constexpr int LPR = BS / EPC;              // lanes sharing one V row
  static constexpr int RPI = 32 / LPR;              // V rows per iteration
  static constexpr int RI = (D + RPI - 1) / RPI;    // V row iterations per lane
  static constexpr bool HALF = std::is_same<T, __half>::value;
  static constexpr bool PACKED = HALF || (std::is_same<T, __nv_bfloat16>::value && !FP8);
  using QS = typename std::conditional<PACKED, T, float>::type;   // q / p storage in smem
  static constexpr int QW = EPC * sizeof(QS) / 16;  // 16-byte words of q (or p) per cache chunk
  static constexpr int TILE_BYTES = D * BS * (int)sizeof(CT);
  static constexpr int Q_BYTES = GT * D * (int)sizeof(QS);
  static constexpr int P_BYTES = GT * BS * (int)sizeof(QS);
  static constexpr bool P_IN_REGS = GT * QW <= 8;
  static_assert(D % EPC == 0 && BS % EPC == 0 && LPR >= 1 && 32 % BS == 0, "geometry");
};

template <typename C>
__host__ __device__ constexpr int warp_smem_bytes(int stages) {
  return stages * 2 * C::TILE_BYTES + ((C::Q_BYTES + 15) / 16) * 16 + ((C::P_BYTES + 15) / 16) * 16 +
         ((stages * 8 + 15) / 16) * 16;
}

// acc += dot(one 16-byte cache chunk, the matching q (or p) words)
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

// ---- the kernel ----------------------------------------------------------------------------------
template <typename T, typename CT, int D, int BS, int GT>
__global__ void __launch_bounds__(128) paged_attention_kernel(const __grid_constant__ Params p) {
  using C = Cfg<T, CT, D, BS, GT>;
  using QS = typename C::QS;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int S = p.stages;
  const int wbytes = warp_smem_bytes<C>(S);
  unsigned char* wbase = smem_raw + (size_t)warp * wbytes;
  const uint32_t tiles_s = smem_u32(wbase);                        // S x {K tile, V tile}
  QS* q_s = reinterpret_cast<QS*>(wbase + S * 2 * C::TILE_BYTES);  // [GT][D]
  QS* p_s = reinterpret_cast<QS*>(reinterpret_cast<unsigned char*>(q_s) +
                                  ((C::Q_BYTES + 15) / 16) * 16);  // [GT][BS]
  const uint32_t bars = smem_u32(reinterpret_cast<unsigned char*>(p_s) +
                                 ((C::P_BYTES + 15) / 16) * 16);   // S mbarriers
  const uint32_t q_sa = smem_u32(q_s), p_sa = smem_u32(p_s);

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // lane geometry
  const int k_tok = lane % BS, k_grp = lane / BS;       // q.k: token, chunk phase
  const int v_row = lane / C::LPR, v_sub = lane % C::LPR;  // p.v: first row, token chunk

  int stage_c = 0;        // consumer cursor over the ring (continues across items)
  uint32_t phase_c = 0;
  int stage_p = 0;        // producer cursor

  const float scale_l2 = p.scale * kLog2e;
  const int total_warps = gridDim.x * (blockDim.x >> 5);

  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&p.counters[0], 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= p.total_items) break;

    // item -> (seq, part, pass, kvh), kvh fastest
    int t = item;
    const int kvh = t % p.num_kv_heads;
    t /= p.num_kv_heads;
    const int pass = t % p.passes;
    t /= p.passes;
    const int part = t % p.n_part;
    const int seq = t / p.n_part;

    const int ctx = p.context_lens[seq];
    const int nblk_ctx = (ctx + BS - 1) / BS;
    const int blk0 = part * p.part_blocks;
    const int blk1 = min(blk0 + p.part_blocks, nblk_ctx);
    const int ntiles = blk1 - blk0;
    const int head0 = kvh * p.group + pass * GT;             // first q head of this pass
    const int nheads = min(GT, p.group - pass * GT);          // live heads in this pass

    if (ntiles <= 0) {
      // empty context: the reference writes zeros (accs stay 0).  Empty partitions of a split
      // launch are never read by the reduce step.
      if (!p.split && part == 0) {
        T* o = reinterpret_cast<T*>(p.out) + ((int64_t)seq * p.num_heads + head0) * D;
        for (int i = lane; i < nheads * D; i += 32) o[i] = ST<T>::from_f(0.f);
      }
      continue;
    }

    const int32_t* bt = p.block_tables + (int64_t)seq * p.max_blocks_per_seq + blk0;
    const char* kbase = reinterpret_cast<const char*>(p.k_cache) + (int64_t)kvh * p.kv_head_stride;
    const char* vbase = reinterpret_cast<const char*>(p.v_cache) + (int64_t)kvh * p.kv_head_stride;

    // block ids: lane i holds the id of tile (batch*32 + i); one coalesced load per 32 tiles
    int bt_cur = lane < ntiles ? bt[lane] : 0;
    int bt_next = 32 + lane < ntiles ? bt[32 + lane] : 0;

    // stage q: [GT][D] as QS (zero rows for dead heads so their dot products stay finite)
    {
      const T* qg = reinterpret_cast<const T*>(p.q) + (int64_t)seq * p.q_stride + (int64_t)head0 * D;
      for (int i = lane; i < GT * D; i += 32) {
        const int h = i / D;
        const float f = h < nheads ? ST<T>::to_f(qg[i]) : 0.f;
        q_s[i] = to_qs<QS>(f);
      }
    }
    __syncwarp();

    // pipeline prologue: put up to S tiles in flight
    const int npro = min(S, ntiles);
    for (int i = 0; i < npro; ++i) {
      const int bid = __shfl_sync(0xffffffffu, bt_cur, i & 31);  // S <= 32
      if (lane == 0) {
        const uint32_t bar = bars + 8 * stage_p;
        const uint32_t dst = tiles_s + stage_p * 2 * C::TILE_BYTES;
        mbar_expect_tx(bar, 2 * C::TILE_BYTES);
        bulk_g2s(dst, kbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES, bar);
        bulk_g2s(dst + C::TILE_BYTES, vbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES, bar);
      }
      stage_p = stage_p + 1 == S ? 0 : stage_p + 1;
    }

    float m[GT], l[GT], acc[GT][C::RI], slope[GT];
#pragma unroll
    for (int h = 0; h < GT; ++h) {
      m[h] = -INFINITY;
      l[h] = 0.f;
      slope[h] = (p.alibi_slopes != nullptr && h < nheads) ? p.alibi_slopes[head0 + h] * kLog2e : 0.f;
#pragma unroll
      for (int i = 0; i < C::RI; ++i) acc[h][i] = 0.f;
    }

    for (int tile = 0; tile < ntiles; ++tile) {
      if (tile > 0 && (tile & 31) == 0) {  // advance the block-id window
        bt_cur = bt_next;
        const int nx = tile + 32 + lane;
        bt_next = nx < ntiles ? bt[nx] : 0;
      }
      const uint32_t ktile = tiles_s + stage_c * 2 * C::TILE_BYTES;
      const uint32_t vtile = ktile + C::TILE_BYTES;
      mbar_wait(bars + 8 * stage_c, phase_c);

      // ---- s = q.k for this lane's token --------------------------------------------------------
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

      // ---- online softmax (log2 domain) ---------------------------------------------------------
      const int tok = (blk0 + tile) * BS + k_tok;
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

      // ---- acc += p.v ---------------------------------------------------------------------------
      const bool last_partial = (blk0 + tile == nblk_ctx - 1) && (ctx % BS != 0);
      const int vvalid = ctx - (blk0 + tile) * BS - v_sub * C::EPC;  // valid tokens in my chunk
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
      __syncwarp();  // every lane is done with this stage (and with p_s)

      // ---- refill the stage just freed ----------------------------------------------------------
      const int nt = tile + S;
      if (nt < ntiles) {
        const int src = ((nt >> 5) == (tile >> 5)) ? bt_cur : bt_next;
        const int bid = __shfl_sync(0xffffffffu, src, nt & 31);
        if (lane == 0) {
          const uint32_t bar = bars + 8 * stage_p;
          const uint32_t dst = tiles_s + stage_p * 2 * C::TILE_BYTES;
          mbar_expect_tx(bar, 2 * C::TILE_BYTES);
          bulk_g2s(dst, kbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES, bar);
          bulk_g2s(dst + C::TILE_BYTES, vbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES,
                   bar);
        }
        stage_p = stage_p + 1 == S ? 0 : stage_p + 1;
      }
      if (++stage_c == S) {
        stage_c = 0;
        phase_c ^= 1;
      }
    }

    // ---- epilogue: finish the reductions, normalise, store -------------------------------------
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
#pragma unroll
    for (int h = 0; h < GT; ++h) {
      if (h < nheads) {
        const int head = head0 + h;
        const float inv = __fdividef(1.f, l[h] + 1e-6f);
        if (!p.split) {
          T* o = reinterpret_cast<T*>(p.out) + ((int64_t)seq * p.num_heads + head) * D;
          if (v_sub == 0) {
#pragma unroll
            for (int i = 0; i < C::RI; ++i) {
              const int row = v_row + i * C::RPI;
              if (D % C::RPI == 0 || row < D) o[row] = ST<T>::from_f(acc[h][i] * inv);
            }
          }
        } else {
          const int64_t pi = ((int64_t)seq * p.num_heads + head) * p.n_part + part;
          if (lane == 0) {
            p.max_logits[pi] = m[h] * kLn2;  // natural-log domain, as the reference stores qk_max
            p.exp_sums[pi] = l[h];
          }
          if (v_sub == 0) {
#pragma unroll
            for (int i = 0; i < C::RI; ++i) {
              const int row = v_row + i * C::RPI;
              if (D % C::RPI == 0 || row < D) {
                const float v = acc[h][i] * inv;
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
    __syncwarp();  // q_s / p_s are rewritten by the next item
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

// ---- merge of partitions (reference: paged_attention_v2_reduce_kernel, :511-602) -----------------
// one warp per (seq, head); partial outputs are normalised per partition, weights are
// exp_sum_j * exp(max_j - max) / (sum + 1e-6)
template <typename T, typename PT, int D>
__global__ void __launch_bounds__(128)
    paged_attention_reduce_kernel(T* __restrict__ out, const float* __restrict__ exp_sums,
                                  const float* __restrict__ max_logits,
                                  const PT* __restrict__ tmp_out,
                                  const int32_t* __restrict__ context_lens, int num_heads,
                                  int n_part, int part_tokens, int total) {
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // seq * num_heads + head
  if (idx >= total) return;
  const int seq = idx / num_heads;
  const int ctx = context_lens[seq];
  int np = (ctx + part_tokens - 1) / part_tokens;
  np = np < n_part ? np : n_part;
  T* o = out + (int64_t)idx * D;
  if (np <= 0) {  // empty context
    for (int i = lane; i < D; i += 32) o[i] = ST<T>::from_f(0.f);
    return;
  }
  const float* ml = max_logits + (int64_t)idx * n_part;
  const float* es = exp_sums + (int64_t)idx * n_part;
  const PT* tp = tmp_out + (int64_t)idx * n_part * D;
  if (np == 1) {  // nothing to merge: copy (reference :523-533)
    for (int i = lane; i < D; i += 32) {
      if constexpr (std::is_same<PT, float>::value)
        o[i] = ST<T>::from_f(tp[i]);
      else
        o[i] = ST<T>::from_f(ST<PT>::to_f(tp[i]));
    }
    return;
  }
  float mx = -INFINITY;
  for (int j = lane; j < np; j += 32) mx = fmaxf(mx, ml[j]);
#pragma unroll
  for (int o2 = 16; o2 >= 1; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
  float sum = 0.f;
  for (int j = lane; j < np; j += 32) sum += es[j] * expf(ml[j] - mx);
#pragma unroll
  for (int o2 = 16; o2 >= 1; o2 >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o2);
  const float inv = __fdividef(1.0f, sum + 1e-6f);
  for (int i = lane; i < D; i += 32) {
    float a = 0.f;
    for (int j = 0; j < np; ++j) {
      const float w = es[j] * expf(ml[j] - mx) * inv;
      float v;
      if constexpr (std::is_same<PT, float>::value)
        v = tp[(int64_t)j * D + i];
      else
        v = ST<PT>::to_f(tp[(int64_t)j * D + i]);
      a += v * w;
    }
    o[i] = ST<T>::from_f(a);
  }
}

}  // namespace attn
}  // namespace vltr
