// attention_inst.cu — explicit instantiations of the paged-attention kernel for ONE data-type mode
// (compiled five times with -DVLTR_ATTN_MODE=0..4 so the modes build in parallel):
//   0: f16 q / f16 cache   1: bf16 / bf16   2: f32 / f32   3: f16 / e5m2   4: bf16|f32 / e5m2
// Head sizes {64,80,96,112,128,256} x block sizes {8,16,32} are the reference's grid
// (csrc/attention/attention_kernels.cu:681-739).
#include <mutex>
#include <unordered_map>

#include "attention_kernel.cuh"

namespace vltr {
namespace attn {

template <typename T, typename CT, int D, int BS, int GT, bool MMA>
static int launch_one(const Params& p0, cudaStream_t stream) {
  using C = Cfg<T, CT, D, BS, GT, MMA>;
  auto kern = paged_attention_kernel<T, CT, D, BS, GT, MMA>;
  // per-instantiation launch plan, computed once: stages, warps/CTA, CTAs/SM
  struct Plan { int stages, warps, smem, ctas_per_sm; bool ok; };
  auto make_plan = [&](int want_warps, int want_stages, int want_ctas) {
    Plan pl{0, 0, 0, 0, false};
    int dev = 0, max_optin = 0, sm_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    // FMA path (one query head per kv head): resident warps matter more than ring depth when
    // the queue is long (measured, tools/attn_tune.py): three 4-warp CTAs with a 2-deep ring
    // beat two with a 3-deep ring by 2-5% on 2-byte caches.
    // Round 2 (profiles/r02_attn_timeline.txt): the tensor-core path is bound by each warp's
    // dependent chain, not by resident warps - with 2-byte caches two CTAs with a 3-deep ring
    // beat three with a 2-deep one by 4-5 % (GQA trace batch 83 -> 79.5 us); the e5m2 variant
    // (registers admit three CTAs) is best with three CTAs and a 3-deep ring.
    const int ctas = want_ctas > 0 ? want_ctas : (MMA ? (C::FP8 ? 3 : 2) : (C::FP8 ? 4 : 3));
    int budget = sm_smem / ctas - 1024 - kPlanBytes;       // 1 KiB per CTA is reserved by the system
    if (budget > max_optin - kPlanBytes) budget = max_optin - kPlanBytes;
    int warps = want_warps > 0 ? want_warps : 4, stages = 0;
    const int top = want_stages > 0 ? want_stages : (MMA && C::FP8 ? 3 : 6);
    // first choice: the wanted CTAs/SM with as many warps as fit; last resort: one CTA per SM
    for (int pass = 0; pass < 2 && stages < 2; ++pass) {
      const int lim = pass == 0 ? budget : max_optin - kPlanBytes;
      for (warps = want_warps > 0 ? want_warps : 4; warps >= 1; warps >>= 1) {
        for (stages = top; stages >= 2; --stages)
          if (warps * warp_smem_bytes<C>(stages) <= lim) break;
        if (stages >= 2) break;
      }
    }
    if (warps < 1 || stages < 2) return pl;
    pl.stages = stages;
    pl.warps = warps;
    pl.smem = warps * warp_smem_bytes<C>(stages) + kPlanBytes;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem) !=
        cudaSuccess) {
      cudaGetLastError();
      return pl;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, pl.smem);
    pl.ctas_per_sm = occ < 1 ? 1 : occ;
    pl.ok = true;
    return pl;
  };
  // two plans: `thr` for long queues (throughput: most resident warps), `lat` for launches whose
  // work items cannot even occupy the resident warps twice (small batches, tensor-parallel
  // shards): there every warp streams one item at the latency-bound rate of its ring, and the
  // deeper ring of two CTAs per SM wins (B=64 GQA: 40 vs 47 us)
  auto set_attr = [&](const Plan& a, const Plan& b) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         a.smem > b.smem ? a.smem : b.smem);
  };
  static Plan thr = make_plan(0, 0, 0), lat = make_plan(0, 0, 2);
  static bool attr_ready = (set_attr(thr, lat), true);
  (void)attr_ready;
  const bool tuned = g_tune.attn_warps > 0 || g_tune.attn_stages > 0 || g_tune.attn_ctas > 0;
  static bool was_tuned = false;
  if (!tuned && was_tuned) {       // back from a development override: restore the attribute
    set_attr(thr, lat);
    was_tuned = false;
  }
  const long long items_ub = (long long)p0.num_seqs * p0.n_part * p0.passes * p0.num_kv_heads;
  Plan plan = (lat.ok && items_ub < 2LL * num_sms() * thr.ctas_per_sm * thr.warps) ? lat : thr;
  if (tuned) {
    plan = make_plan((int)g_tune.attn_warps, (int)g_tune.attn_stages, (int)g_tune.attn_ctas);
    was_tuned = true;
  }
  if (!plan.ok) return fail(VLTR_ERR_CUDA, "paged_attention: cannot configure shared memory");
  Params p = p0;
  p.stages = plan.stages;
  int grid = num_sms() * plan.ctas_per_sm;
  const long long need = (items_ub + plan.warps - 1) / plan.warps;
  if (grid > need) grid = (int)need;
  if (grid < 1) grid = 1;
  p.trace = reinterpret_cast<long long*>(g_tune.attn_trace);
  launch(kern, dim3(grid), dim3(plan.warps * 32), (size_t)plan.smem, stream, p);
  VLTR_LAUNCH_CHECK("paged_attention");
  return VLTR_OK;
}

// largest head-group tile of the FMA path that keeps the accumulators in registers (GT * RI <= 64)
template <typename T, typename CT, int D, int BS>
static constexpr int gt_cap() {
  constexpr int RI = Cfg<T, CT, D, BS, 1, false>::RI;
  return RI <= 8 ? 8 : (RI <= 16 ? 4 : (RI <= 32 ? 2 : 1));
}

template <typename T, typename CT, int D, int BS>
static int launch_gt(Params& p, cudaStream_t stream) {
  const int G = p.group;
  if constexpr (sizeof(T) == 2) {
    if (G >= 2) {   // tensor-core path
      const int gt = G <= 8 ? 8 : 16;
      p.passes = (G + gt - 1) / gt;
      if (gt == 8) return launch_one<T, CT, D, BS, 8, true>(p, stream);
      return launch_one<T, CT, D, BS, 16, true>(p, stream);
    }
    if constexpr (BS % (16 / (int)sizeof(CT)) != 0) {
      // e5m2 cache with block size 8: a V row is half a 16-byte chunk; the FMA path cannot tile
      // it, the tensor-core path can (G == 1 rides row 0 of the tile)
      p.passes = 1;
      return launch_one<T, CT, D, BS, 8, true>(p, stream);
    } else {
      p.passes = 1;
      return launch_one<T, CT, D, BS, 1, false>(p, stream);
    }
  } else {
    if constexpr (BS % (16 / (int)sizeof(CT)) != 0) {
      return fail(VLTR_ERR_UNSUPPORTED,
                  "Unsupported block size: %d for fp32 queries with an fp8 kv cache", BS);
    } else {
      constexpr int CAP = gt_cap<T, CT, D, BS>();
      int gt = G <= 1 ? 1 : (G <= 2 ? 2 : (G <= 4 ? 4 : 8));
      if (gt > CAP) gt = CAP;
      p.passes = (G + gt - 1) / gt;
      switch (gt) {
        case 1: return launch_one<T, CT, D, BS, 1, false>(p, stream);
        case 2:
          if constexpr (CAP >= 2) return launch_one<T, CT, D, BS, 2, false>(p, stream);
          break;
        case 4:
          if constexpr (CAP >= 4) return launch_one<T, CT, D, BS, 4, false>(p, stream);
          break;
        case 8:
          if constexpr (CAP >= 8) return launch_one<T, CT, D, BS, 8, false>(p, stream);
          break;
      }
      return fail(VLTR_ERR_UNSUPPORTED, "paged_attention: no kernel for group tile %d", gt);
    }
  }
}

template <typename T, typename CT, int BS>
static int launch_d(Params& p, int head_size, cudaStream_t stream) {
  switch (head_size) {
    case 64: return launch_gt<T, CT, 64, BS>(p, stream);
    case 80: return launch_gt<T, CT, 80, BS>(p, stream);
    case 96: return launch_gt<T, CT, 96, BS>(p, stream);
    case 112: return launch_gt<T, CT, 112, BS>(p, stream);
    case 128: return launch_gt<T, CT, 128, BS>(p, stream);
    case 256: return launch_gt<T, CT, 256, BS>(p, stream);
    default: return fail(VLTR_ERR_UNSUPPORTED, "Unsupported head size: %d", head_size);
  }
}

template <typename T, typename CT>
static int launch_bs(Params& p, int head_size, int block_size, cudaStream_t stream) {
  switch (block_size) {
    case 8: return launch_d<T, CT, 8>(p, head_size, stream);
    case 16: return launch_d<T, CT, 16>(p, head_size, stream);
    case 32: return launch_d<T, CT, 32>(p, head_size, stream);
    default: return fail(VLTR_ERR_UNSUPPORTED, "Unsupported block size: %d", block_size);
  }
}

#if VLTR_ATTN_MODE == 0
int launch_mode_f16(Params& p, int head_size, int block_size, cudaStream_t s) {
  return launch_bs<__half, __half>(p, head_size, block_size, s);
}
#elif VLTR_ATTN_MODE == 1
int launch_mode_bf16(Params& p, int head_size, int block_size, cudaStream_t s) {
  return launch_bs<__nv_bfloat16, __nv_bfloat16>(p, head_size, block_size, s);
}
#elif VLTR_ATTN_MODE == 2
int launch_mode_f32(Params& p, int head_size, int block_size, cudaStream_t s) {
  return launch_bs<float, float>(p, head_size, block_size, s);
}
#elif VLTR_ATTN_MODE == 3
int launch_mode_f16_fp8(Params& p, int head_size, int block_size, cudaStream_t s) {
  return launch_bs<__half, uint8_t>(p, head_size, block_size, s);
}
#elif VLTR_ATTN_MODE == 4
int launch_mode_bf16_fp8(Params& p, int head_size, int block_size, cudaStream_t s) {
  return launch_bs<__nv_bfloat16, uint8_t>(p, head_size, block_size, s);
}
int launch_mode_f32_fp8(Params& p, int head_size, int block_size, cudaStream_t s) {
  return launch_bs<float, uint8_t>(p, head_size, block_size, s);
}
#endif

}  // namespace attn
}  // namespace vltr
