// prefill_attention.cu — attention of prefill chunks over the paged KV cache ("prefix prefill").
//
// Replaces the Triton kernel the reference runs for EVERY prefill chunk when chunked prefill is on
// (vllm/attention/ops/prefix_prefill.py:10-755 `context_attention_fwd`, reached from
// PagedAttention.forward_prefix, vllm/attention/ops/paged_attn.py:163-195, called at
// vllm/attention/backends/flash_attn.py:237-252): for every new token i of sequence b,
//     out[i] = softmax(scale * q[i] . [K_cache(b, 0..ctx_b) ; k_new(b, 0..i)]^T) . [V_cache ; v_new]
// i.e. each query attends to the cached prefix and, causally, to the chunk's own tokens.
//
// B200 shape.  The engine writes the chunk's K/V into the paged cache before attention runs
// (write_to_paged_cache, flash_attn.py:192), so a chunk token is a decode query whose context is
// ctx_b + i + 1 tokens of that cache.  The kernel is the decode kernel's tensor-core path with
// the 16 rows of the mma tile holding (token, query head) pairs that share a kv head:
//   rows r = (token r / G, head r % G), G = query heads per kv head, 16/G tokens per item;
//   K/V tiles by TMA bulk copies into a per-warp mbarrier ring, ldmatrix fragments straight from
//   the paged layout, online softmax on the fragments with a per-ROW causal limit.
// When the caller cannot guarantee that the cache already holds the chunk (the reference's unit
// test of this op does not), `scatter_new_kv` first writes k/v into their slots - in the engine
// flow that rewrites identical bytes.  f16 / bf16 models (fp32 is not served through this path).
#include <mutex>

#include "attention_kernel.cuh"

namespace vltr {
namespace prefill {

using namespace vltr::attn;

struct PParams {
  void* out;                    // [T, Hq, D], row stride o_stride (elements)
  const void* q;                // [T, Hq, D], row stride q_stride
  const void* k_cache;
  const void* v_cache;
  const int32_t* block_tables;  // [B, max_blocks_per_seq]
  const int32_t* start_loc;     // [B] first row of the sequence's chunk in q / out
  const int32_t* seq_lens;      // [B] context + chunk
  const int32_t* ctx_lens;      // [B] tokens already cached before this chunk
  const float* alibi_slopes;    // [Hq] or null
  int num_seqs, num_heads, num_kv_heads, group, tq;   // tq = chunk tokens per item (16 / G)
  int max_blocks_per_seq;
  float scale;
  int64_t q_stride, o_stride;   // elements
  int64_t kv_block_stride, kv_head_stride;   // bytes
  int stages;
  int* counters;                // [0] next item, [1] finished warps (self-resetting)
};

constexpr int kMaxSeqs = 2047;  // prefix-sum plan in shared memory

template <typename T, typename CT, int D, int BS>
__global__ void __launch_bounds__(128) prefill_attention_kernel(const __grid_constant__ PParams p) {
  using C = Cfg<T, CT, D, BS, 16, true>;
  using MT = typename C::MT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int S = p.stages;
  const int wbytes = warp_smem_bytes<C>(S);
  unsigned char* wbase = smem_raw + (size_t)warp * wbytes;
  const uint32_t tiles_s = smem_u32(wbase);
  const uint32_t bars = tiles_s + S * 2 * C::TILE_BYTES;
  int* cum = reinterpret_cast<int*>(smem_raw + (size_t)nwarps * wbytes);   // [num_seqs + 1]
  const int g = lane >> 2, t4 = lane & 3;
  const int G = p.group, TQ = p.tq;

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // plan: exclusive prefix sums of the row groups per sequence
  for (int s = threadIdx.x; s < p.num_seqs; s += blockDim.x) {
    const int n = p.seq_lens[s] - p.ctx_lens[s];
    cum[s + 1] = n > 0 ? (n + TQ - 1) / TQ : 0;
  }
  __syncthreads();
  if (warp == 0) {
    int carry = 0;
    for (int base = 0; base < p.num_seqs; base += 32) {
      const int s = base + lane;
      const int n = s < p.num_seqs ? cum[s + 1] : 0;
      int incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      __syncwarp();
      if (s < p.num_seqs) cum[s + 1] = carry + incl;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) cum[0] = 0;
  }
  __syncthreads();
  const int total_items = cum[p.num_seqs] * p.num_kv_heads;
  const float scale_l2 = p.scale * kLog2e;
  const int total_warps = gridDim.x * nwarps;

  int stage_c = 0, stage_p = 0;
  uint32_t phase_c = 0;

  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&p.counters[0], 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total_items) break;
    const int kvh = item % p.num_kv_heads;
    const int t = item / p.num_kv_heads;
    int lo = 0, hi = p.num_seqs;          // largest s with cum[s] <= t
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] <= t) lo = mid; else hi = mid;
    }
    const int seq = lo, rg = t - cum[lo];
    const int ctx0 = p.ctx_lens[seq];
    const int n_new = p.seq_lens[seq] - ctx0;
    const int tok_q0 = rg * TQ;                         // first chunk token of this item
    const int nq = min(TQ, n_new - tok_q0);             // live chunk tokens
    const int row0 = p.start_loc[seq] + tok_q0;         // their first row in q / out
    const int kmax = ctx0 + tok_q0 + nq;                // keys visible to the last live token
    const int ntiles = (kmax + BS - 1) / BS;
    const int32_t* bt = p.block_tables + (int64_t)seq * p.max_blocks_per_seq;
    const char* kbase = reinterpret_cast<const char*>(p.k_cache) + (int64_t)kvh * p.kv_head_stride;
    const char* vbase = reinterpret_cast<const char*>(p.v_cache) + (int64_t)kvh * p.kv_head_stride;

    // per-row (token, head) mapping and causal limit; dead rows keep limit 1 (finite softmax)
    int lim[2], tokq[2], hd[2];
    bool live[2];
    float slope[2], m[2], l[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = g + 8 * r;
      tokq[r] = row / G;
      hd[r] = row % G;
      live[r] = tokq[r] < nq && row < TQ * G;
      lim[r] = live[r] ? ctx0 + tok_q0 + tokq[r] + 1 : 1;
      slope[r] = (p.alibi_slopes != nullptr && live[r]) ? p.alibi_slopes[kvh * G + hd[r]] * kLog2e : 0.f;
      m[r] = -INFINITY;
      l[r] = 0.f;
    }
    float o_acc[C::ND][4];
#pragma unroll
    for (int i = 0; i < C::ND; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
    uint32_t qa[C::KS][4];
#pragma unroll
    for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        uint32_t qlo = 0, qhi = 0;
        if (live[r]) {
          const T* src = reinterpret_cast<const T*>(p.q) + (int64_t)(row0 + tokq[r]) * p.q_stride +
                         (int64_t)(kvh * G + hd[r]) * D + ks * 16;
          if constexpr (C::FP8) {
            const uint2 w = *reinterpret_cast<const uint2*>(src + 4 * t4);
            if constexpr (std::is_same<T, __half>::value) {
              qlo = w.x;
              qhi = w.y;
            } else {
              const float2 a = unpack2<__nv_bfloat16>(w.x), b = unpack2<__nv_bfloat16>(w.y);
              qlo = pack2<__half>(a.x, a.y);
              qhi = pack2<__half>(b.x, b.y);
            }
          } else {
            qlo = *reinterpret_cast<const uint32_t*>(src + 2 * t4);
            qhi = *reinterpret_cast<const uint32_t*>(src + 8 + 2 * t4);
          }
        }
        qa[ks][r] = qlo;
        qa[ks][2 + r] = qhi;
      }
    }

    // block ids: lane i holds the id of tile 32 * win + i; the next window is fetched a window
    // ahead so the table read never sits on the issue path
    int win = 0;
    int bt_cur = lane < ntiles ? bt[lane] : 0;
    int bt_nxt = 32 + lane < ntiles ? bt[32 + lane] : 0;
    auto issue = [&](int tile) {
      if ((tile >> 5) != win) {
        win = tile >> 5;
        bt_cur = bt_nxt;
        const int nx = 32 * (win + 1) + lane;
        bt_nxt = nx < ntiles ? bt[nx] : 0;
      }
      const int bid = __shfl_sync(0xffffffffu, bt_cur, tile & 31);
      if (lane == 0) {
        const uint32_t bar = bars + 8 * stage_p;
        const uint32_t dst = tiles_s + stage_p * 2 * C::TILE_BYTES;
        mbar_expect_tx(bar, 2 * C::TILE_BYTES);
        bulk_g2s(dst, kbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES, bar);
        bulk_g2s(dst + C::TILE_BYTES, vbase + (int64_t)bid * p.kv_block_stride, C::TILE_BYTES, bar);
      }
      stage_p = stage_p + 1 == S ? 0 : stage_p + 1;
    };
    const int npro = min(S, ntiles);
    for (int i = 0; i < npro; ++i) issue(i);

    for (int tile = 0; tile < ntiles; ++tile) {
      const uint32_t ktile = tiles_s + stage_c * 2 * C::TILE_BYTES;
      const uint32_t vtile = ktile + C::TILE_BYTES;
      mbar_wait(bars + 8 * stage_c, phase_c);
      const int tok0 = tile * BS;
      if (tok0 + BS > kmax) {
        // zero the V columns past the last visible key (they may hold NaNs, p = 0 there)
        const int valid = kmax - tok0;
        unsigned char* vt = wbase + (size_t)stage_c * 2 * C::TILE_BYTES + C::TILE_BYTES;
        for (int i = lane; i < D * BS; i += 32)
          if (i % BS >= valid) reinterpret_cast<CT*>(vt)[i] = CT(0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
      }
      float s[C::NT][4];
#pragma unroll
      for (int nt = 0; nt < C::NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if constexpr (!C::FP8) {
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) {
          if constexpr (C::NT >= 2) {
#pragma unroll
            for (int np = 0; np < C::NT / 2; ++np) {
              uint32_t b[4];
              const int j = lane >> 3;
              ldsm_x4(ktile + (((2 * ks + (j & 1)) * BS + 16 * np + 8 * (j >> 1) + (lane & 7)) << 4), b);
              mma16816<MT>(s[2 * np], qa[ks], b[0], b[1]);
              mma16816<MT>(s[2 * np + 1], qa[ks], b[2], b[3]);
            }
          } else {
            uint32_t b[4];
            const int j = (lane >> 3) & 1;
            ldsm_x4(ktile + (((2 * ks + j) * BS + (lane & 7)) << 4), b);
            mma16816<MT>(s[0], qa[ks], b[0], b[1]);
          }
        }
      } else {
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) {
            const uint32_t w = lds4(ktile + ((ks * BS + 8 * nt + g) << 4) + 4 * t4);
            mma16816<MT>(s[nt], qa[ks], __byte_perm(w, 0, 0x1404), __byte_perm(w, 0, 0x3424));
          }
        }
      }
      uint32_t pa[C::KT][4];
#pragma unroll
      for (int kt = 0; kt < C::KT; ++kt) pa[kt][0] = pa[kt][1] = pa[kt][2] = pa[kt][3] = 0u;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int tok = tok0 + 8 * nt + 2 * t4 + e;
            float sv = s[nt][2 * r + e] * scale_l2 + slope[r] * (float)(tok - lim[r] + 1);
            sv = tok < lim[r] ? sv : -INFINITY;
            s[nt][2 * r + e] = sv;
            mx = fmaxf(mx, sv);
          }
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        // a tile past this row's causal limit contributes nothing (mx = -inf): keep the state
        const float m_new = fmaxf(m[r], mx);
        const float alpha = m_new == -INFINITY ? 1.f : exp2f(m[r] - m_new);
        const float m_use = m_new == -INFINITY ? 0.f : m_new;
        m[r] = m_new;
        l[r] *= alpha;
#pragma unroll
        for (int i = 0; i < C::ND; ++i) {
          o_acc[i][2 * r] *= alpha;
          o_acc[i][2 * r + 1] *= alpha;
        }
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) {
          const uint32_t pk = pack2<MT>(exp2f(s[nt][2 * r] - m_use), exp2f(s[nt][2 * r + 1] - m_use));
          const float2 pf = unpack2<MT>(pk);
          l[r] += pf.x + pf.y;
          pa[nt >> 1][(nt & 1) * 2 + r] = pk;
        }
      }
      if constexpr (!C::FP8) {
#pragma unroll
        for (int kt = 0; kt < C::KT; ++kt) {
#pragma unroll
          for (int nd = 0; nd < C::ND; nd += 2) {
            uint32_t b[4];
            const int j = lane >> 3;
            if constexpr (BS >= 16) {
              ldsm_x4(vtile + ((8 * (nd + (j >> 1)) + (lane & 7)) * BS + 16 * kt + 8 * (j & 1)) * 2, b);
              mma16816<MT>(o_acc[nd], pa[kt], b[0], b[1]);
              mma16816<MT>(o_acc[nd + 1], pa[kt], b[2], b[3]);
            } else {
              ldsm_x4(vtile + ((8 * (nd + (j >> 1)) + (lane & 7)) * BS) * 2, b);
              mma16816<MT>(o_acc[nd], pa[kt], b[0], 0u);
              mma16816<MT>(o_acc[nd + 1], pa[kt], b[2], 0u);
            }
          }
        }
      } else {
#pragma unroll
        for (int kt = 0; kt < C::KT; ++kt) {
#pragma unroll
          for (int nd = 0; nd < C::ND; ++nd) {
            const uint32_t row = vtile + (8 * nd + g) * BS + 16 * kt;
            uint32_t w0, w1 = 0;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w0) : "r"(row + 2 * t4));
            if constexpr (BS >= 16)
              asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w1) : "r"(row + 8 + 2 * t4));
            mma16816<MT>(o_acc[nd], pa[kt], __byte_perm(w0, 0, 0x1404), __byte_perm(w1, 0, 0x1404));
          }
        }
      }
      __syncwarp();
      if (++stage_c == S) {
        stage_c = 0;
        phase_c ^= 1;
      }
      if (tile + S < ntiles) issue(tile + S);
    }

    // epilogue
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
      if (live[r]) {
        const float inv = __fdividef(1.f, l[r]);
        T* o = reinterpret_cast<T*>(p.out) + (int64_t)(row0 + tokq[r]) * p.o_stride +
               (int64_t)(kvh * G + hd[r]) * D;
#pragma unroll
        for (int nd = 0; nd < C::ND; ++nd)
          *reinterpret_cast<uint32_t*>(o + 8 * nd + 2 * t4) =
              pack2<T>(o_acc[nd][2 * r] * inv, o_acc[nd][2 * r + 1] * inv);
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    const int done = atomicAdd(&p.counters[1], 1);
    if (done == total_warps - 1) {
      p.counters[0] = 0;
      p.counters[1] = 0;
      __threadfence();
    }
  }
}

// writes the chunk's k / v rows into their cache slots (what write_to_paged_cache does in the
// engine; needed when this op is called on its own).  One warp per (row, kv head).
template <typename T, typename CT, bool FP8>
__global__ void __launch_bounds__(256)
    scatter_new_kv_kernel(const T* __restrict__ key, const T* __restrict__ value,
                          CT* __restrict__ k_cache, CT* __restrict__ v_cache,
                          const int32_t* __restrict__ block_tables,
                          const int32_t* __restrict__ start_loc, const int32_t* __restrict__ seq_lens,
                          const int32_t* __restrict__ ctx_lens, int num_seqs, int max_blocks,
                          int num_kv_heads, int head_size, int block_size, int64_t k_stride,
                          int64_t v_stride, int64_t kv_block_elems) {
  constexpr int XE = 16 / sizeof(CT);
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.y;
  const int n_new = seq_lens[seq] - ctx_lens[seq];
  const int rows = n_new * num_kv_heads;
  for (int rh = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rh < rows;
       rh += gridDim.x * (blockDim.x >> 5)) {
    const int i = rh / num_kv_heads, head = rh % num_kv_heads;
    const int pos = ctx_lens[seq] + i;
    const int64_t blk = block_tables[(int64_t)seq * max_blocks + pos / block_size];
    const int off = pos % block_size;
    const T* ks = key + (int64_t)(start_loc[seq] + i) * k_stride + (int64_t)head * head_size;
    const T* vs = value + (int64_t)(start_loc[seq] + i) * v_stride + (int64_t)head * head_size;
    CT* kd = k_cache + blk * kv_block_elems + (int64_t)head * head_size * block_size;
    CT* vd = v_cache + blk * kv_block_elems + (int64_t)head * head_size * block_size;
    for (int d = lane; d < head_size; d += 32) {
      CT kv, vv;
      if constexpr (FP8) {
        kv = to_fp8<T>(ks[d]);
        vv = to_fp8<T>(vs[d]);
      } else {
        kv = ks[d];
        vv = vs[d];
      }
      kd[((d / XE) * block_size + off) * XE + d % XE] = kv;
      vd[(int64_t)d * block_size + off] = vv;
    }
  }
}

static int* g_counters[64] = {nullptr};
static std::mutex g_mu;

template <typename T, typename CT, int D, int BS>
static int launch_one(PParams& p, int total_rows_ub, cudaStream_t stream) {
  using C = Cfg<T, CT, D, BS, 16, true>;
  auto kern = prefill_attention_kernel<T, CT, D, BS>;
  struct Plan { int stages, warps, smem, ctas_per_sm; bool ok; };
  constexpr int kPlan = (kMaxSeqs + 1) * 4;
  static Plan plan = [&]() {
    Plan pl{0, 0, 0, 0, false};
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int budget = (max_optin - 2048) / 2 - kPlan;
    int warps = 4, stages = 0;
    for (; warps >= 1; warps >>= 1) {
      const int lim = warps == 4 ? budget : max_optin - kPlan;
      for (stages = 6; stages >= 2; --stages)
        if (warps * warp_smem_bytes<C>(stages) <= lim) break;
      if (stages >= 2) break;
    }
    if (warps < 1 || stages < 2) return pl;
    pl.stages = stages;
    pl.warps = warps;
    pl.smem = warps * warp_smem_bytes<C>(stages) + kPlan;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem) != cudaSuccess) {
      cudaGetLastError();
      return pl;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, pl.smem);
    pl.ctas_per_sm = occ < 1 ? 1 : occ;
    pl.ok = true;
    return pl;
  }();
  if (!plan.ok) return fail(VLTR_ERR_CUDA, "context_attention: cannot configure shared memory");
  p.stages = plan.stages;
  int grid = num_sms() * plan.ctas_per_sm;
  const long long items_ub = ((long long)total_rows_ub / p.tq + p.num_seqs) * p.num_kv_heads;
  const long long need = (items_ub + plan.warps - 1) / plan.warps;
  if (grid > need) grid = (int)need;
  if (grid < 1) grid = 1;
  kern<<<grid, plan.warps * 32, plan.smem, stream>>>(p);
  VLTR_LAUNCH_CHECK("context_attention_fwd");
  return VLTR_OK;
}

template <typename T, typename CT>
static int launch_bs_d(PParams& p, int head_size, int block_size, int rows, cudaStream_t s) {
#define VLTR_PF(DD, BB) return launch_one<T, CT, DD, BB>(p, rows, s);
#define VLTR_PF_D(BB)                                                                   \
  switch (head_size) {                                                                  \
    case 64: VLTR_PF(64, BB) case 80: VLTR_PF(80, BB) case 96: VLTR_PF(96, BB)           \
    case 112: VLTR_PF(112, BB) case 128: VLTR_PF(128, BB) case 256: VLTR_PF(256, BB)     \
    default: return fail(VLTR_ERR_UNSUPPORTED, "Unsupported head size: %d", head_size);  \
  }
  switch (block_size) {
    case 8: VLTR_PF_D(8)
    case 16: VLTR_PF_D(16)
    case 32: VLTR_PF_D(32)
    default: return fail(VLTR_ERR_UNSUPPORTED, "Unsupported block size: %d", block_size);
  }
#undef VLTR_PF_D
#undef VLTR_PF
}

}  // namespace prefill
}  // namespace vltr

extern "C" int vltr_context_attention_fwd(
    void* out, const void* query, const void* key, const void* value, void* key_cache,
    void* value_cache, const int32_t* block_tables, const int32_t* start_loc,
    const int32_t* seq_lens, const int32_t* ctx_lens, int num_seqs, int num_tokens, int num_heads,
    int num_kv_heads, int head_size, int block_size, int max_num_blocks_per_seq, float scale,
    const float* alibi_slopes, int64_t q_stride, int64_t k_stride, int64_t v_stride,
    int64_t o_stride, int64_t kv_block_stride, int64_t kv_head_stride, int dtype, int kv_fp8,
    int cache_has_new_tokens, void* stream) {
  using namespace vltr;
  using namespace vltr::prefill;
  VLTR_CHECK(num_seqs >= 0 && num_tokens >= 0 && num_heads > 0 && num_kv_heads > 0,
             VLTR_ERR_INVALID, "context_attention: bad sizes");
  VLTR_CHECK(num_heads % num_kv_heads == 0, VLTR_ERR_INVALID,
             "context_attention: num_heads %d not divisible by num_kv_heads %d", num_heads,
             num_kv_heads);
  VLTR_CHECK(dtype == VLTR_F16 || dtype == VLTR_BF16, VLTR_ERR_UNSUPPORTED,
             "context_attention: Unsupported data type: %d (f16 / bf16)", dtype);
  const int G = num_heads / num_kv_heads;
  VLTR_CHECK(G <= 16, VLTR_ERR_UNSUPPORTED,
             "context_attention: %d query heads per kv head (max 16)", G);
  VLTR_CHECK(num_seqs <= kMaxSeqs, VLTR_ERR_UNSUPPORTED, "context_attention: %d sequences (max %d)",
             num_seqs, kMaxSeqs);
  if (num_seqs == 0 || num_tokens == 0) return VLTR_OK;
  const int cs = kv_fp8 ? 1 : 2;
  VLTR_CHECK(kv_head_stride == (int64_t)head_size * block_size, VLTR_ERR_INVALID,
             "context_attention: a (block, head) KV tile must be contiguous");
  VLTR_CHECK(((uintptr_t)key_cache | (uintptr_t)value_cache | (uintptr_t)query | (uintptr_t)out) % 16 == 0 &&
                 (kv_block_stride * cs) % 16 == 0 && (q_stride * 2) % 16 == 0 && (o_stride * 2) % 4 == 0,
             VLTR_ERR_INVALID, "context_attention: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0;
  VLTR_CUDA(cudaGetDevice(&dev));
  VLTR_CHECK(dev >= 0 && dev < 64, VLTR_ERR_INVALID, "device index %d out of range", dev);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_counters[dev] == nullptr) {
      VLTR_CUDA(cudaMalloc(&g_counters[dev], sizeof(int) * 2 * 16));
      VLTR_CUDA(cudaMemset(g_counters[dev], 0, sizeof(int) * 2 * 16));
    }
  }
  static unsigned slot = 0;
  if (!cache_has_new_tokens) {
    VLTR_CHECK(key != nullptr && value != nullptr, VLTR_ERR_INVALID,
               "context_attention: key / value needed to fill the cache");
    const dim3 grid(64, num_seqs);
#define VLTR_SC(TT)                                                                             \
  if (kv_fp8)                                                                                   \
    scatter_new_kv_kernel<TT, uint8_t, true><<<grid, 256, 0, s>>>(                              \
        (const TT*)key, (const TT*)value, (uint8_t*)key_cache, (uint8_t*)value_cache,           \
        block_tables, start_loc, seq_lens, ctx_lens, num_seqs, max_num_blocks_per_seq,          \
        num_kv_heads, head_size, block_size, k_stride, v_stride, kv_block_stride);              \
  else                                                                                          \
    scatter_new_kv_kernel<TT, TT, false><<<grid, 256, 0, s>>>(                                  \
        (const TT*)key, (const TT*)value, (TT*)key_cache, (TT*)value_cache, block_tables,       \
        start_loc, seq_lens, ctx_lens, num_seqs, max_num_blocks_per_seq, num_kv_heads,          \
        head_size, block_size, k_stride, v_stride, kv_block_stride);
    if (dtype == VLTR_F16) { VLTR_SC(__half) } else { VLTR_SC(__nv_bfloat16) }
#undef VLTR_SC
    VLTR_LAUNCH_CHECK("context_attention_scatter");
  }
  PParams p{};
  p.out = out;
  p.q = query;
  p.k_cache = key_cache;
  p.v_cache = value_cache;
  p.block_tables = block_tables;
  p.start_loc = start_loc;
  p.seq_lens = seq_lens;
  p.ctx_lens = ctx_lens;
  p.alibi_slopes = alibi_slopes;
  p.num_seqs = num_seqs;
  p.num_heads = num_heads;
  p.num_kv_heads = num_kv_heads;
  p.group = G;
  p.tq = 16 / G;
  p.max_blocks_per_seq = max_num_blocks_per_seq;
  p.scale = scale;
  p.q_stride = q_stride;
  p.o_stride = o_stride;
  p.kv_block_stride = kv_block_stride * cs;
  p.kv_head_stride = kv_head_stride * cs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    p.counters = g_counters[dev] + 2 * (slot++ % 16);
  }
  if (dtype == VLTR_F16)
    return kv_fp8 ? launch_bs_d<__half, uint8_t>(p, head_size, block_size, num_tokens, s)
                  : launch_bs_d<__half, __half>(p, head_size, block_size, num_tokens, s);
  return kv_fp8 ? launch_bs_d<__nv_bfloat16, uint8_t>(p, head_size, block_size, num_tokens, s)
                : launch_bs_d<__nv_bfloat16, __nv_bfloat16>(p, head_size, block_size, num_tokens, s);
}
