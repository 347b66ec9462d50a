// opt_predictor.cu — forward pass of the learning-to-rank predictor (OPT-125M / OPT-350M
// sequence classifier), prefill only, no KV cache.
//
// Replaces what AUXLLMEngine.obtain_aux_scores (vllm/engine/aux_llm_engine.py:332-412) runs on
// the GPU through a whole second engine: OPTForSequenceClassification.forward + compute_logits
// (vllm/model_executor/models/opt.py:362-409; decoder :181-310, layer :101-178, attention :57-98,
// flash_attn_varlen prefill path attention/backends/flash_attn.py:217-235).
//
// Op order and rounding points follow the reference (every op writes T = fp16/bf16 activations,
// fp32 accumulation inside):
//   h = embed_tokens[id] (-> project_in) + embed_positions[pos + 2]
//   per layer (pre-LN, OPT-125M):  h += out_proj(attn(qkv_proj(LN1(h))));  h += fc2(relu(fc1(LN2(h))))
//   per layer (post-LN, OPT-350M): h = LN1(h + out_proj(attn(qkv_proj(h))));  h = LN2(h + fc2(relu(fc1(h))))
//   last token of every prompt -> final LN -> (project_out) -> logits = x . score.weight^T
//
// B200 shape
//   * the dense layers (the only contraction in scope) run on the 5th-generation tensor cores:
//     one tcgen05.mma stream per CTA issued by a single thread, A/B tiles staged in shared memory
//     by TMA (128-byte swizzle) through a 4-stage mbarrier ring, fp32 accumulators in TMEM, and a
//     tcgen05.ld epilogue that fuses bias, ReLU and the residual add before the one global store;
//   * LayerNorm is one warp per token row, attention is a causal varlen kernel with one thread
//     per query row (online softmax in registers, K/V tiles broadcast from shared memory) - the
//     predictor sees a few hundred tokens per tick, these are latency-, not throughput-, critical;
//   * the whole forward is ~8 launches per layer on one stream with no host synchronisation; the
//     caller reads the [num_prompts, num_labels] logits back once.
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace vltr {
namespace opt {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, one 128 x N x 16 step
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// dense layer on tcgen05:  D[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) + resid[M,N]
// ------------------------------------------------------------------------------------------------
constexpr int kBM = 128;      // UMMA M (one TMEM lane per output row)
constexpr int kBK = 64;       // 64 halfs = 128 bytes = one swizzle row
constexpr int kStages = 3;   // 3 x 32 KB (BN = 128): two CTAs per SM, one's epilogue under the other's MMAs
constexpr int kGemmThreads = 128;

template <int BN>
__host__ __device__ constexpr int gemm_smem_bytes() {
  return kStages * (kBM * kBK * 2 + BN * kBK * 2) + 1024 /*alignment slack*/ + 256 /*barriers*/;
}

// K-major, 128-byte swizzled operand tile: rows are 128 B apart, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout SWIZZLE_128B = 2 [61,64))
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// cute::UMMA::InstrDescriptor for kind::f16: c_format F32 (1) [4,6), a/b format [7,10)/[10,13)
// (0 = f16, 1 = bf16), both K-major, N>>3 [17,23), M>>4 [24,29)
template <typename T, int BN>
__device__ __forceinline__ uint32_t make_idesc() {
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(kBM >> 4) << 24);
}

template <typename T, int BN, bool RELU>
__global__ void __launch_bounds__(kGemmThreads)
    gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a,
                   const __grid_constant__ CUtensorMap map_b, T* __restrict__ D,
                   const T* __restrict__ bias, const T* __restrict__ resid, int M, int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = kBM * kBK * 2, B_BYTES = BN * kBK * 2;
  constexpr uint32_t STAGE = A_BYTES + B_BYTES;
  const uint32_t bars = base + kStages * STAGE;       // full[kStages], empty[kStages], tmem_full
  const uint32_t full0 = bars, empty0 = bars + 8 * kStages, tfull = bars + 16 * kStages;
  const uint32_t slot = tfull + 8;
  volatile uint32_t* slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int nk = (K + kBK - 1) / kBK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(slot, BN);   // BN fp32 accumulator columns (power of two >= 32)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  // barriers and the TMEM allocation above touch no global memory: under programmatic dependent
  // launch they overlap the tail of the previous kernel of the forward pass
  pdl_wait();
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    // ---- TMA producer ---------------------------------------------------------------------------
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(empty0 + 8 * s, ph ^ 1);
      mbar_expect_tx(full0 + 8 * s, STAGE);
      tma_load_2d(base + s * STAGE, &map_a, kb * kBK, m0, full0 + 8 * s);
      tma_load_2d(base + s * STAGE + A_BYTES, &map_b, kb * kBK, n0, full0 + 8 * s);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer -----------------------------------------------------------------------------
    const uint32_t idesc = make_idesc<T, BN>();
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      const uint64_t ad = make_desc_sw128(base + s * STAGE);
      const uint64_t bd = make_desc_sw128(base + s * STAGE + A_BYTES);
#pragma unroll
      for (int k = 0; k < kBK / 16; ++k)   // +32 bytes (2 x 16 B) per K = 16 step inside the atom
        umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
      umma_commit(empty0 + 8 * s);          // frees the stage when these MMAs have read it
    }
    umma_commit(tfull);                      // accumulator complete
  }

  // ---- epilogue: TMEM -> registers -> bias / relu / residual -> global ---------------------------
  mbar_wait(tfull, 0);
  __syncwarp();
  tc_fence_after();
  const int row = m0 + warp * 32 + lane;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(trow + c * 32, r);
    const int col = n0 + c * 32;
    if (row < M) {
      T* drow = D + (int64_t)row * N + col;
      const T* rrow = resid != nullptr ? resid + (int64_t)row * N + col : nullptr;
#pragma unroll
      for (int v = 0; v < 4; ++v) {          // 4 x 16-byte stores of 8 elements
        if (col + v * 8 >= N) break;         // last column tile of an N that is not a multiple of BN
        alignas(16) T o[8];
        alignas(16) T rs[8];
        alignas(16) T bs[8];
        if (rrow != nullptr) *reinterpret_cast<uint4*>(rs) = *reinterpret_cast<const uint4*>(rrow + v * 8);
        if (bias != nullptr) *reinterpret_cast<uint4*>(bs) = __ldg(reinterpret_cast<const uint4*>(bias + col + v * 8));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float x = __uint_as_float(r[v * 8 + e]);
          if (bias != nullptr) x += ST<T>::to_f(bs[e]);
          if (RELU) x = fmaxf(x, 0.f);
          if (rrow != nullptr) x = rt<T>(x) + ST<T>::to_f(rs[e]);   // linear output is T, then + residual
          o[e] = ST<T>::from_f(x);
        }
        st_16(drow + v * 8, *reinterpret_cast<uint4*>(o));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, BN);
}

// ------------------------------------------------------------------------------------------------
// dense layer, persistent form (the default): one CTA per SM walks a strided list of output tiles
// with three specialised roles, so the epilogue of tile i runs under the MMAs of tile i+1:
//   warp 0      TMA producer   - A / W k-blocks into a kStages2-deep shared-memory ring
//   warp 1      MMA issuer     - tcgen05.mma into one of TWO TMEM accumulators (2 x BN columns)
//   warps 2..5  epilogue       - tcgen05.ld their 32-lane quarter, bias / ReLU in registers, round
//                                to T, transpose through a padded shared-memory slab so that the
//                                residual read and the store are full 128-byte rows (4 rows per
//                                warp instruction) instead of one 16-byte piece per row
// Barriers: full/empty per ring stage (TMA <-> MMA), tmem_full/tmem_empty per accumulator
// (MMA <-> epilogue).  K for these layers is only 768..3072 (12..48 k-blocks), so without the
// overlap the epilogue is as long as the main loop.
// ------------------------------------------------------------------------------------------------
constexpr int kGemm2Threads = 192;
constexpr int kSlabRow = 144;                       // 128 B of T + 16 B pad: conflict-free both ways
template <int BN>
__host__ __device__ constexpr int gemm2_stages() { return BN == 128 ? 5 : 6; }
template <int BN>
__host__ __device__ constexpr int gemm2_smem_bytes() {
  return gemm2_stages<BN>() * (kBM * kBK * 2 + BN * kBK * 2) + 4 * 32 * kSlabRow + 1024 + 256;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <typename T, int BN, bool RELU>
__global__ void __launch_bounds__(kGemm2Threads, 1)
    gemm_tn_persistent_kernel(const __grid_constant__ CUtensorMap map_a,
                              const __grid_constant__ CUtensorMap map_b, T* __restrict__ D,
                              const T* __restrict__ bias, const T* __restrict__ resid, int M, int N,
                              int K) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int S = gemm2_stages<BN>();
  constexpr uint32_t A_BYTES = kBM * kBK * 2, B_BYTES = BN * kBK * 2, STAGE = A_BYTES + B_BYTES;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slab0 = base + S * STAGE;
  const uint32_t bars = slab0 + 4 * 32 * kSlabRow;
  const uint32_t full0 = bars, empty0 = bars + 8 * S, tfull0 = bars + 16 * S,
                 tempty0 = tfull0 + 16, slot = tempty0 + 16;
  unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));     // generic view of `base`
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // N need not be a multiple of BN (vocabulary shards of the lm_head): TMA zero-fills the W rows
  // past N, the epilogue skips the columns past N
  const int num_n = (N + BN - 1) / BN, num_m = (M + kBM - 1) / kBM, total = num_m * num_n;
  const int nk = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 4);      // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        // m fastest: the CTAs working at the same time share W tiles (the lm_head at 256 rows has
        // two M tiles per 128 vocabulary rows - with n fastest the 262 MB weight matrix would
        // stream from HBM twice)
        const int m0 = (tile % num_m) * kBM, n0 = (tile / num_m) * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full0 + 8 * s, STAGE);
          tma_load_2d(base + s * STAGE, &map_a, kb * kBK, m0, full0 + 8 * s);
          tma_load_2d(base + s * STAGE + A_BYTES, &map_b, kb * kBK, n0, full0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc<T, BN>();
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++t) {
        const int buf = t & 1;
        mbar_wait(tempty0 + 8 * buf, ((t >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(buf * BN);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint64_t ad = make_desc_sw128(base + s * STAGE);
          const uint64_t bd = make_desc_sw128(base + s * STAGE + A_BYTES);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_f16(acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull0 + 8 * buf);
      }
    }
  } else {
    // ---- epilogue warps: TMEM lanes [32q, 32q+32) are the only ones warp (id % 4 == q) may read --
    const int q = warp & 3;
    const uint32_t slab = slab0 + (uint32_t)((warp - 2) * 32 * kSlabRow);
    const unsigned char* slab_gen = gen + (slab - base);
    int t = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++t) {
      const int m0 = (tile % num_m) * kBM, n0 = (tile / num_m) * BN;
      const int buf = t & 1;
      mbar_wait(tfull0 + 8 * buf, (t >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem + (uint32_t)(buf * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        // my row's 64 columns -> bias / relu -> T -> slab row (128 B)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld32(trow + c * 64 + h * 32, r);
          const int col = n0 + c * 64 + h * 32;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            alignas(16) T o[8];
            alignas(16) T bs[8];
            const bool have_bias = bias != nullptr && col + v * 8 < N;
            if (have_bias) *reinterpret_cast<uint4*>(bs) = __ldg(reinterpret_cast<const uint4*>(bias + col + v * 8));
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float x = __uint_as_float(r[v * 8 + e]);
              if (have_bias) x += ST<T>::to_f(bs[e]);
              if (RELU) x = fmaxf(x, 0.f);
              o[e] = ST<T>::from_f(x);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(slab + lane * kSlabRow + h * 64 + v * 16),
                         "r"(reinterpret_cast<uint4*>(o)->x), "r"(reinterpret_cast<uint4*>(o)->y),
                         "r"(reinterpret_cast<uint4*>(o)->z), "r"(reinterpret_cast<uint4*>(o)->w)
                         : "memory");
          }
        }
        __syncwarp();
        // transposed pass: 8 lanes cover one 128-byte row, 4 rows per instruction
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), ch = lane & 7;
          const int row = m0 + q * 32 + rr;
          if (row < M && n0 + c * 64 + ch * 8 < N) {
            uint4 val = *reinterpret_cast<const uint4*>(slab_gen + rr * kSlabRow + ch * 16);
            const int64_t off = (int64_t)row * N + n0 + c * 64 + ch * 8;
            if (resid != nullptr) {
              alignas(16) T rs[8];
              alignas(16) T o[8];
              *reinterpret_cast<uint4*>(rs) = *reinterpret_cast<const uint4*>(resid + off);
              const T* lin = reinterpret_cast<const T*>(&val);
#pragma unroll
              for (int e = 0; e < 8; ++e)    // linear output already rounded to T, then + residual
                o[e] = ST<T>::from_f(ST<T>::to_f(lin[e]) + ST<T>::to_f(rs[e]));
              val = *reinterpret_cast<uint4*>(o);
            }
            st_16(D + off, val);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 2 * BN);
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
// out[t, :] = tok[ids[t], :E] (+ pos[positions[t] + 2, :] when ADD_POS; requires E == H then)
template <typename T>
__global__ void __launch_bounds__(128)
    embed_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ positions,
                 const T* __restrict__ tok, const T* __restrict__ pos, T* __restrict__ out, int E,
                 int add_pos) {
  pdl_wait();
  pdl_trigger();
  const int t = blockIdx.x;
  const T* src = tok + (int64_t)ids[t] * E;
  const T* p = pos + ((int64_t)positions[t] + 2) * E;   // OPTLearnedPositionalEmbedding offset 2
  for (int i = threadIdx.x * 8; i < E; i += blockDim.x * 8) {
    alignas(16) T a[8];
    *reinterpret_cast<uint4*>(a) = *reinterpret_cast<const uint4*>(src + i);
    if (add_pos) {
      alignas(16) T b[8];
      *reinterpret_cast<uint4*>(b) = *reinterpret_cast<const uint4*>(p + i);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = ST<T>::from_f(ST<T>::to_f(a[e]) + ST<T>::to_f(b[e]));
    }
    *reinterpret_cast<uint4*>(out + (int64_t)t * E + i) = *reinterpret_cast<uint4*>(a);
  }
}

// h[t, :] += pos[positions[t] + 2, :]   (after project_in, OPT-350M)
template <typename T>
__global__ void __launch_bounds__(128)
    add_pos_kernel(const int32_t* __restrict__ positions, const T* __restrict__ pos,
                   T* __restrict__ h, int H) {
  pdl_wait();
  pdl_trigger();
  const int t = blockIdx.x;
  const T* p = pos + ((int64_t)positions[t] + 2) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    h[(int64_t)t * H + i] = ST<T>::from_f(ST<T>::to_f(h[(int64_t)t * H + i]) + ST<T>::to_f(p[i]));
}

// LayerNorm, one warp per row.  rows == nullptr: row r of `in` -> row r of `out`;
// else row rows[r] of `in` -> row r of `out` (gather of the last token of every prompt).
template <typename T>
__global__ void __launch_bounds__(128)
    layernorm_kernel(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ w,
                     const T* __restrict__ b, const int32_t* __restrict__ rows, int nrows, int H,
                     float eps, int identity) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const T* x = in + (int64_t)(rows != nullptr ? rows[r] : r) * H;
  T* y = out + (int64_t)r * H;
  if (identity) {   // no final layer norm: plain gather
    for (int i = lane; i < H; i += 32) y[i] = x[i];
    return;
  }
  // H is a multiple of 64: every lane owns H/256 (+1) 16-byte chunks, kept in registers
  constexpr int kMaxChunks = 8;                 // H <= 2048
  const int nchunk = H / 8;
  uint4 v[kMaxChunks];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      v[c] = *reinterpret_cast<const uint4*>(x + ch * 8);
      const T* e = reinterpret_cast<const T*>(&v[c]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += ST<T>::to_f(e[k]);
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  const float mean = s / H;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    if (lane + 32 * c < nchunk) {
      const T* e = reinterpret_cast<const T*>(&v[c]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = ST<T>::to_f(e[k]) - mean;
        var += d * d;
      }
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) var += __shfl_xor_sync(0xffffffffu, var, m);
  const float rstd = rsqrtf(var / H + eps);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      const T* e = reinterpret_cast<const T*>(&v[c]);
      alignas(16) T wv[8], bv[8], o[8];
      *reinterpret_cast<uint4*>(wv) = __ldg(reinterpret_cast<const uint4*>(w + ch * 8));
      *reinterpret_cast<uint4*>(bv) = __ldg(reinterpret_cast<const uint4*>(b + ch * 8));
#pragma unroll
      for (int k = 0; k < 8; ++k)
        o[k] = ST<T>::from_f((ST<T>::to_f(e[k]) - mean) * rstd * ST<T>::to_f(wv[k]) + ST<T>::to_f(bv[k]));
      *reinterpret_cast<uint4*>(y + ch * 8) = *reinterpret_cast<uint4*>(o);
    }
  }
}

// causal varlen attention, head size 64.  qkv [M, 3H] (q | k | v), out [M, H].
//
// Flash-style tile kernel on the warp-level tensor-core path (mma.sync m16n8k16, fp32 accumulate):
// grid (query tiles of 64 rows over max_len - heaviest first, heads, prompts), 4 warps x 16 query
// rows.  K/V tiles of 64 keys x 64 dims are staged by cp.async into a double-buffered, XOR-
// swizzled shared-memory ring; Q.K^T takes its B fragments with ldmatrix, P.V with
// ldmatrix.trans; the online softmax lives in registers (row statistics reduced across the 4
// lanes that share a row), P is rounded to T before P.V as flash-attn does (the reference's
// prefill path, attention/backends/flash_attn.py:217-235).  The contraction is tiny next to the
// dense layers (head size 64), so it stays on mma.sync rather than paying a TMEM round trip.
constexpr int kAttnRows = 64, kAttnKeys = 64, kHD = 64, kAttnThreads = 128;

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;          // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                          uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&d)[4], const uint32_t (&a)[4],
                                                  uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4],
                                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// tile row r (128 bytes = 8 chunks of 16 B), chunk c: XOR swizzle keeps ldmatrix conflict-free
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
    attention_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                     const int32_t* __restrict__ cu_seqlens, int H, float scale) {
  // [Q | K0 | V0 | K1 | V1], each 64 rows x 128 B
  __shared__ __align__(128) unsigned char smem[5 * kAttnKeys * 128];
  pdl_wait();
  pdl_trigger();
  const int prompt = blockIdx.z, head = blockIdx.y;
  const int t0 = cu_seqlens[prompt], len = cu_seqlens[prompt + 1] - t0;
  const int q0 = ((int)gridDim.x - 1 - (int)blockIdx.x) * kAttnRows;   // long (late) tiles first
  if (q0 >= len) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ld = 3 * (int64_t)H;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sq = sbase;
  const T* base = qkv + (int64_t)t0 * ld + head * kHD;

  // ---- stage Q (group 0) and the first K/V tile (group 1) --------------------------------------
  for (int i = tid; i < kAttnRows * 8; i += kAttnThreads) {
    const int r = i >> 3, c = i & 7;
    cp_async_16(sq + tile_off(r, c), base + (int64_t)min(q0 + r, len - 1) * ld + c * 8, q0 + r < len);
  }
  cp_async_commit();
  const int kend = min(len, q0 + kAttnRows);          // causal: keys up to this tile's last query
  const int ntiles = (kend + kAttnKeys - 1) / kAttnKeys;
  auto stage_kv = [&](int kt, int buf) {
    const uint32_t sk = sbase + (1 + 2 * buf) * (kAttnKeys * 128), sv = sk + kAttnKeys * 128;
    const int k0 = kt * kAttnKeys;
    for (int i = tid; i < kAttnKeys * 8; i += kAttnThreads) {
      const int r = i >> 3, c = i & 7;
      const bool ok = k0 + r < len;
      const T* kp = base + (int64_t)min(k0 + r, len - 1) * ld + H + c * 8;
      cp_async_16(sk + tile_off(r, c), kp, ok);
      cp_async_16(sv + tile_off(r, c), kp + H, ok);      // rows past the prompt are zero-filled
    }
  };
  stage_kv(0, 0);
  cp_async_commit();

  // ---- Q fragments: 16 rows x 64 dims per warp = 4 k-steps of A (m16 x k16) --------------------
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[4][4];
  {
    const int mi = lane >> 3, r = lane & 7;
    const int row = warp * 16 + (mi & 1) * 8 + r;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(sq + tile_off(row, ks * 2 + (mi >> 1)), qf[ks]);
  }
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g + 8 of this warp
  const float sl2 = scale * 1.4426950408889634f;
  const int g = lane >> 2, tq = lane & 3;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;        // query positions in the prompt

  for (int kt = 0; kt < ntiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) stage_kv(kt + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sk = sbase + (1 + 2 * buf) * (kAttnKeys * 128), sv = sk + kAttnKeys * 128;
    const int k0 = kt * kAttnKeys;
    // ---- S = Q . K^T  (16 x 64 per warp) --------------------------------------------------------
    float sc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[n][e] = 0.f;
      const int krow = n * 8 + (lane & 7), mi = lane >> 3;
      uint32_t b[4];
      ldsm_x4(sk + tile_off(krow, mi), b);           // dims 0..31: k-steps 0 and 1
      mma_16816<T>(sc[n], qf[0], b[0], b[1]);
      mma_16816<T>(sc[n], qf[1], b[2], b[3]);
      ldsm_x4(sk + tile_off(krow, 4 + mi), b);       // dims 32..63: k-steps 2 and 3
      mma_16816<T>(sc[n], qf[2], b[0], b[1]);
      mma_16816<T>(sc[n], qf[3], b[2], b[3]);
    }
    // ---- scale, causal mask (diagonal tile only), online softmax ------------------------------
    const bool diag = k0 + kAttnKeys > q0 + warp * 16;      // some key of the tile is past a row
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int key = k0 + n * 8 + tq * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = sc[n][e] * sl2;
        if (diag && key + (e & 1) > ((e & 2) ? row1 : row0)) v = -INFINITY;
        sc[n][e] = v;
      }
      mx0 = fmaxf(mx0, fmaxf(sc[n][0], sc[n][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[n][2], sc[n][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    // key 0 is visible to every row, so mx is finite from the first tile on
    const float a0 = exp2f(m0 - mx0), a1 = exp2f(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    l0 *= a0;
    l1 *= a1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= a0;
      o[n][1] *= a0;
      o[n][2] *= a1;
      o[n][3] *= a1;
    }
    uint32_t pf[4][4];        // P as A fragments: k-step j covers keys 16j .. 16j+15
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float p0 = exp2f(sc[n][0] - m0), p1 = exp2f(sc[n][1] - m0);
      const float p2 = exp2f(sc[n][2] - m1), p3 = exp2f(sc[n][3] - m1);
      l0 += p0 + p1;
      l1 += p2 + p3;
      pf[n >> 1][(n & 1) * 2 + 0] = pack2<T>(p0, p1);
      pf[n >> 1][(n & 1) * 2 + 1] = pack2<T>(p2, p3);
    }
    // ---- O += P . V ---------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mi = lane >> 3, r = lane & 7;
      const int vrow = j * 16 + (mi & 1) * 8 + r;
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        ldsm_x4_t(sv + tile_off(vrow, np * 2 + (mi >> 1)), b);
        mma_16816<T>(o[np * 2], pf[j], b[0], b[1]);
        mma_16816<T>(o[np * 2 + 1], pf[j], b[2], b[3]);
      }
    }
    __syncthreads();          // everyone is done with `buf` before it is refilled
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  T* op = out + (int64_t)t0 * H + head * kHD + tq * 2;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (row0 < len)
      *reinterpret_cast<uint32_t*>(op + (int64_t)row0 * H + n * 8) = pack2<T>(o[n][0] * i0, o[n][1] * i0);
    if (row1 < len)
      *reinterpret_cast<uint32_t*>(op + (int64_t)row1 * H + n * 8) = pack2<T>(o[n][2] * i1, o[n][3] * i1);
  }
}

// logits[i, l] = (T)(x[i, :] . w[l, :]) as float;  one warp per (prompt, label)
template <typename T>
__global__ void __launch_bounds__(128)
    score_kernel(const T* __restrict__ x, const T* __restrict__ w, float* __restrict__ logits,
                 int n, int labels, int E) {
  pdl_wait();
  pdl_trigger();
  const int idx = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= n * labels) return;
  const int i = idx / labels, lbl = idx % labels;
  float s = 0.f;
  for (int e = lane; e < E; e += 32)
    s += ST<T>::to_f(x[(int64_t)i * E + e]) * ST<T>::to_f(w[(int64_t)lbl * E + e]);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) logits[idx] = rt<T>(s);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// 2-D map over a row-major [rows, cols] 16-bit matrix, box = [box_rows, 64 cols], 128 B swizzle
static int make_map(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows, int dtype) {
  EncodeTiledFn fn = encode_fn();
  VLTR_CHECK(fn != nullptr, VLTR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, dtype == VLTR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                        2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLTR_CHECK(r == CUDA_SUCCESS, VLTR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return VLTR_OK;
}

template <typename T, int BN, bool RELU>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, void* D, const void* bias,
                       const void* resid, int M, int N, int K, cudaStream_t stream) {
  if (g_tune.opt_gemm == 1) {       // the one-tile-per-CTA form (A/B knob, tools/predictor_bench.py)
    auto kern = gemm_tn_kernel<T, BN, RELU>;
    static bool configured = false;
    if (!configured) {
      VLTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     gemm_smem_bytes<BN>()));
      configured = true;
    }
    dim3 grid((N + BN - 1) / BN, (M + kBM - 1) / kBM);
    launch(kern, grid, dim3(kGemmThreads), (size_t)gemm_smem_bytes<BN>(), stream, ma, mb, (T*)D,
           (const T*)bias, (const T*)resid, M, N, K);
    VLTR_LAUNCH_CHECK("opt_gemm");
    return VLTR_OK;
  }
  auto kern = gemm_tn_persistent_kernel<T, BN, RELU>;
  static bool configured = false;
  if (!configured) {
    VLTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   gemm2_smem_bytes<BN>()));
    configured = true;
  }
  const int total = ((N + BN - 1) / BN) * ((M + kBM - 1) / kBM);
  const int ctas = total < num_sms() ? total : num_sms();
  launch(kern, dim3(ctas), dim3(kGemm2Threads), (size_t)gemm2_smem_bytes<BN>(), stream, ma, mb,
         (T*)D, (const T*)bias, (const T*)resid, M, N, K);
  VLTR_LAUNCH_CHECK("opt_gemm");
  return VLTR_OK;
}

// D = act(A . W^T + bias) + resid;  A [M,K], W [N,K], all row-major T
static int gemm(const void* A, const void* W, void* D, const void* bias, const void* resid,
                int M, int N, int K, int relu, int dtype, cudaStream_t stream) {
  VLTR_CHECK(M > 0 && N > 0 && K > 0, VLTR_ERR_INVALID, "opt_gemm: empty problem");
  VLTR_CHECK(N % 8 == 0 && K % 8 == 0, VLTR_ERR_UNSUPPORTED,
             "opt_gemm: N %d and K %d must be multiples of 8", N, K);
  VLTR_CHECK(((uintptr_t)A | (uintptr_t)W | (uintptr_t)D | (uintptr_t)bias | (uintptr_t)resid) % 16 == 0,
             VLTR_ERR_INVALID, "opt_gemm: operands must be 16-byte aligned");
  // wide tiles when they still give every SM a CTA, else narrow ones
  const int tiles128 = ((N + 127) / 128) * ((M + kBM - 1) / kBM);
  // (narrow tiles for few rows to fill the last wave were tried for the lm_head - 1000 instead of
  // 500 tiles for 32000 x 4096 at 256 rows - and measured slower, 106 vs 84 us: every tile
  // re-reads its A rows, profiles/r02_lm_head.txt)
  const bool wide = tiles128 >= num_sms();
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M, K, kBM, dtype);
  if (rc) return rc;
  rc = make_map(&mb, W, N, K, wide ? 128 : 64, dtype);
  if (rc) return rc;
#define VLTR_GEMM(TT)                                                                            \
  if (wide)                                                                                      \
    return relu ? launch_gemm<TT, 128, true>(ma, mb, D, bias, resid, M, N, K, stream)            \
                : launch_gemm<TT, 128, false>(ma, mb, D, bias, resid, M, N, K, stream);          \
  return relu ? launch_gemm<TT, 64, true>(ma, mb, D, bias, resid, M, N, K, stream)               \
              : launch_gemm<TT, 64, false>(ma, mb, D, bias, resid, M, N, K, stream);
  if (dtype == VLTR_F16) {
    VLTR_GEMM(__half)
  } else if (dtype == VLTR_BF16) {
    VLTR_GEMM(__nv_bfloat16)
  }
#undef VLTR_GEMM
  return fail(VLTR_ERR_UNSUPPORTED, "opt_gemm: dtype %d (f16 / bf16 only)", dtype);
}

struct Model {
  vltr_opt_config cfg;
  vltr_opt_weights w;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <typename T>
static int forward_t(const Model& mdl, const int32_t* ids, const int32_t* positions,
                     const int32_t* cu_seqlens, const int32_t* last_rows, int M, int P, int max_len,
                     char* ws, float* logits, cudaStream_t s) {
  const vltr_opt_config& c = mdl.cfg;
  const vltr_opt_weights& w = mdl.w;
  const int H = c.hidden, E = c.embed_dim, F = c.ffn, dt = c.dtype;
  const size_t es = sizeof(T);
  // workspace carve-up
  T* h = (T*)ws;            ws += align256((size_t)M * H * es);
  T* x = (T*)ws;            ws += align256((size_t)M * H * es);
  T* qkv = (T*)ws;          ws += align256((size_t)M * 3 * H * es);
  T* att = (T*)ws;          ws += align256((size_t)M * H * es);
  T* ff = (T*)ws;           ws += align256((size_t)M * F * es);
  T* emb = (T*)ws;          ws += align256((size_t)M * E * es);
  T* last = (T*)ws;         ws += align256((size_t)P * H * es);
  T* last_e = (T*)ws;       ws += align256((size_t)P * E * es);
  int rc;
  const float scale = 1.0f / sqrtf((float)(H / c.heads));

  if (E == H) {
    launch(embed_kernel<T>, dim3(M), dim3(128), 0, s, ids, positions, (const T*)w.embed_tokens,
                                      (const T*)w.embed_positions, h, E, 1);
    VLTR_LAUNCH_CHECK("opt_embed");
  } else {
    launch(embed_kernel<T>, dim3(M), dim3(128), 0, s, ids, positions, (const T*)w.embed_tokens,
                                      (const T*)w.embed_positions, emb, E, 0);
    VLTR_LAUNCH_CHECK("opt_embed");
    if ((rc = gemm(emb, w.project_in, h, nullptr, nullptr, M, H, E, 0, dt, s))) return rc;
    launch(add_pos_kernel<T>, dim3(M), dim3(128), 0, s, positions, (const T*)w.embed_positions, h, H);
    VLTR_LAUNCH_CHECK("opt_add_pos");
  }
  const int ln_grid = (M + 3) / 4;
  const dim3 agrid((max_len + kAttnRows - 1) / kAttnRows, c.heads, P);
  for (int L = 0; L < c.layers; ++L) {
    const T* a_in = h;
    if (c.do_layer_norm_before) {
      launch(layernorm_kernel<T>, dim3(ln_grid), dim3(128), 0, s, h, x, (const T*)w.ln1_w[L], (const T*)w.ln1_b[L],
                                                  nullptr, M, H, 1e-5f, 0);
      VLTR_LAUNCH_CHECK("opt_layernorm");
      a_in = x;
    }
    if ((rc = gemm(a_in, w.qkv_w[L], qkv, w.qkv_b[L], nullptr, M, 3 * H, H, 0, dt, s))) return rc;
    launch(attention_kernel<T>, dim3(agrid), dim3(kAttnThreads), 0, s, qkv, att, cu_seqlens, H, scale);
    VLTR_LAUNCH_CHECK("opt_attention");
    if (c.do_layer_norm_before) {
      // h = h + out_proj(att)
      if ((rc = gemm(att, w.out_w[L], h, w.out_b[L], h, M, H, H, 0, dt, s))) return rc;
      launch(layernorm_kernel<T>, dim3(ln_grid), dim3(128), 0, s, h, x, (const T*)w.ln2_w[L], (const T*)w.ln2_b[L],
                                                  nullptr, M, H, 1e-5f, 0);
      VLTR_LAUNCH_CHECK("opt_layernorm");
      if ((rc = gemm(x, w.fc1_w[L], ff, w.fc1_b[L], nullptr, M, F, H, 1, dt, s))) return rc;
      if ((rc = gemm(ff, w.fc2_w[L], h, w.fc2_b[L], h, M, H, F, 0, dt, s))) return rc;
    } else {
      // h = LN1(h + out_proj(att)); h = LN2(h + fc2(relu(fc1(h))))
      if ((rc = gemm(att, w.out_w[L], x, w.out_b[L], h, M, H, H, 0, dt, s))) return rc;
      launch(layernorm_kernel<T>, dim3(ln_grid), dim3(128), 0, s, x, h, (const T*)w.ln1_w[L], (const T*)w.ln1_b[L],
                                                  nullptr, M, H, 1e-5f, 0);
      VLTR_LAUNCH_CHECK("opt_layernorm");
      if ((rc = gemm(h, w.fc1_w[L], ff, w.fc1_b[L], nullptr, M, F, H, 1, dt, s))) return rc;
      if ((rc = gemm(ff, w.fc2_w[L], x, w.fc2_b[L], h, M, H, F, 0, dt, s))) return rc;
      launch(layernorm_kernel<T>, dim3(ln_grid), dim3(128), 0, s, x, h, (const T*)w.ln2_w[L], (const T*)w.ln2_b[L],
                                                  nullptr, M, H, 1e-5f, 0);
      VLTR_LAUNCH_CHECK("opt_layernorm");
    }
  }
  // last token of every prompt -> final LN (row-wise, so only those rows are normalised)
  launch(layernorm_kernel<T>, dim3((P + 3) / 4), dim3(128), 0, s, h, last, (const T*)w.final_ln_w,
                                                  (const T*)w.final_ln_b, last_rows, P, H, 1e-5f,
                                                  c.final_layer_norm ? 0 : 1);
  VLTR_LAUNCH_CHECK("opt_final_layernorm");
  const T* feat = last;
  if (E != H) {
    if ((rc = gemm(last, w.project_out, last_e, nullptr, nullptr, P, E, H, 0, dt, s))) return rc;
    feat = last_e;
  }
  launch(score_kernel<T>, dim3((P * c.num_labels + 3) / 4), dim3(128), 0, s, feat, (const T*)w.score, logits, P,
                                                             c.num_labels, E);
  VLTR_LAUNCH_CHECK("opt_score");
  return VLTR_OK;
}

}  // namespace opt
}  // namespace vltr

extern "C" int vltr_opt_create(const vltr_opt_config* cfg, const vltr_opt_weights* weights,
                               void** handle) {
  using namespace vltr;
  VLTR_CHECK(cfg != nullptr && weights != nullptr && handle != nullptr, VLTR_ERR_INVALID,
             "opt_create: null argument");
  VLTR_CHECK(cfg->layers > 0 && cfg->layers <= VLTR_OPT_MAX_LAYERS, VLTR_ERR_UNSUPPORTED,
             "opt_create: %d layers (max %d)", cfg->layers, VLTR_OPT_MAX_LAYERS);
  VLTR_CHECK(cfg->heads > 0 && cfg->hidden == cfg->heads * 64, VLTR_ERR_UNSUPPORTED,
             "opt_create: head size must be 64 (hidden %d, heads %d)", cfg->hidden, cfg->heads);
  VLTR_CHECK(cfg->hidden <= 2048, VLTR_ERR_UNSUPPORTED,
             "opt_create: hidden %d (LayerNorm keeps a row in registers: max 2048)", cfg->hidden);
  VLTR_CHECK(cfg->hidden % 64 == 0 && cfg->ffn % 64 == 0 && cfg->embed_dim % 64 == 0,
             VLTR_ERR_UNSUPPORTED, "opt_create: hidden / ffn / embed_dim must be multiples of 64");
  VLTR_CHECK(cfg->dtype == VLTR_F16 || cfg->dtype == VLTR_BF16, VLTR_ERR_UNSUPPORTED,
             "opt_create: dtype %d (f16 / bf16 only)", cfg->dtype);
  VLTR_CHECK(cfg->num_labels >= 1, VLTR_ERR_INVALID, "opt_create: num_labels %d", cfg->num_labels);
  opt::Model* m = new opt::Model{*cfg, *weights};
  *handle = m;
  return VLTR_OK;
}

extern "C" void vltr_opt_destroy(void* handle) { delete reinterpret_cast<vltr::opt::Model*>(handle); }

extern "C" size_t vltr_opt_workspace_bytes(const void* handle, int max_tokens, int max_prompts) {
  using namespace vltr::opt;
  const Model* m = reinterpret_cast<const Model*>(handle);
  const vltr_opt_config& c = m->cfg;
  const size_t es = 2, M = max_tokens, P = max_prompts;
  return align256(M * c.hidden * es) * 3 + align256(M * 3 * c.hidden * es) +
         align256(M * c.ffn * es) + align256(M * c.embed_dim * es) + align256(P * c.hidden * es) +
         align256(P * c.embed_dim * es) + 1024;
}

extern "C" int vltr_opt_forward(const void* handle, const int32_t* token_ids,
                                const int32_t* positions, const int32_t* cu_seqlens,
                                const int32_t* last_rows, int num_tokens, int num_prompts,
                                int max_len, void* workspace, size_t workspace_bytes,
                                float* logits, void* stream) {
  using namespace vltr;
  VLTR_CHECK(handle != nullptr, VLTR_ERR_INVALID, "opt_forward: null handle");
  const opt::Model* m = reinterpret_cast<const opt::Model*>(handle);
  VLTR_CHECK(num_tokens >= 0 && num_prompts >= 0 && max_len >= 0, VLTR_ERR_INVALID,
             "opt_forward: negative size");
  if (num_tokens == 0 || num_prompts == 0) return VLTR_OK;
  VLTR_CHECK(max_len <= m->cfg.max_pos, VLTR_ERR_INVALID,
             "opt_forward: prompt of %d tokens exceeds max_position_embeddings %d", max_len,
             m->cfg.max_pos);
  VLTR_CHECK(workspace != nullptr &&
                 workspace_bytes >= vltr_opt_workspace_bytes(handle, num_tokens, num_prompts),
             VLTR_ERR_INVALID, "opt_forward: workspace too small");
  VLTR_CHECK((uintptr_t)workspace % 256 == 0, VLTR_ERR_INVALID,
             "opt_forward: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (m->cfg.dtype == VLTR_F16)
    return opt::forward_t<__half>(*m, token_ids, positions, cu_seqlens, last_rows, num_tokens,
                                  num_prompts, max_len, (char*)workspace, logits, s);
  return opt::forward_t<__nv_bfloat16>(*m, token_ids, positions, cu_seqlens, last_rows, num_tokens,
                                       num_prompts, max_len, (char*)workspace, logits, s);
}

extern "C" int vltr_opt_gemm(const void* A, const void* W, void* D, const void* bias,
                             const void* resid, int M, int N, int K, int relu, int dtype,
                             void* stream) {
  return vltr::opt::gemm(A, W, D, bias, resid, M, N, K, relu, dtype, (cudaStream_t)stream);
}

// logits = hidden . lm_head^T (+ bias): LogitsProcessor._get_logits
// (vllm/model_executor/layers/logits_processor.py:61-64) on the same tcgen05 dense layer.  The
// vocabulary shard of a tensor-parallel lm_head (32000 / 8 = 4000 rows) is not a multiple of the
// tile width: the last column tile is partial.
extern "C" int vltr_lm_head_logits(const void* hidden, const void* weight, const void* bias,
                                   void* logits, int num_rows, int vocab_rows, int hidden_size,
                                   int dtype, void* stream) {
  using namespace vltr;
  VLTR_CHECK(num_rows >= 0 && vocab_rows >= 0 && hidden_size > 0, VLTR_ERR_INVALID,
             "lm_head_logits: negative size");
  if (num_rows == 0 || vocab_rows == 0) return VLTR_OK;
  return opt::gemm(hidden, weight, logits, bias, nullptr, num_rows, vocab_rows, hidden_size, 0,
                   dtype, (cudaStream_t)stream);
}
