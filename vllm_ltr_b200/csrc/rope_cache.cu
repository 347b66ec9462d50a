// rope_cache.cu — rotary embedding of a step's queries and keys FUSED with the KV-cache append.
//
// In the reference these are two ops with a round trip of the keys through HBM between them:
//   q, k = self.rotary_emb(positions, q, k)            vllm/model_executor/models/llama.py:159
//       -> ops.rotary_embedding                         csrc/pos_encoding_kernels.cu:124-172
//   PagedAttention.write_to_paged_cache(k, v, ...)     vllm/attention/backends/flash_attn.py:192
//       -> cache_ops.reshape_and_cache                  csrc/cache_kernels.cu:235-276
// Both are launch-latency-bound at decode (256 tokens): together 6-7 us of a ~100 us Llama-3-8B
// layer.  The attention seam (vllm_ltr_b200/attention_backend.py) sees queries, keys, values,
// positions, the slot mapping and the cache together, so one launch does both:
//   * query rows: rotated in place, (token, head, 16-byte chunk pair) flattened over threads;
//   * key rows: one warp per (token, kv head) - the warp of the FIRST token of a run of
//     consecutive slots (every prefill chunk; a decode token is a run of one) rotates the run's
//     keys, writes them back in place (the caller's `key` holds the rotated values afterwards,
//     as after the reference's op) and stores them into the paged K layout, 16 bytes per
//     (token, chunk), consecutive lanes on consecutive tokens of a chunk;
//   * value rows: the same warp copies the run's values into the paged V layout (16-byte stores
//     where the run covers an aligned group of a block row, element stores otherwise).
// Same arithmetic as the separate kernels ("in T" rotation, e5m2 saturate-to-finite): results
// are bit-identical to rotary_embedding followed by reshape_and_cache (tests/test_gpu_rope_cache.py).
// NeoX style with rot_dim == head_size and 16-byte-aligned operands only - anything else is
// VLTR_ERR_UNSUPPORTED and the caller issues the two ops.
#include "common.cuh"

namespace vltr {

template <int BYTES>
__device__ __forceinline__ void store_bytes(void* dst, const void* src) {
  if constexpr (BYTES == 16)
    st_16(dst, *reinterpret_cast<const uint4*>(src));
  else if constexpr (BYTES == 8)
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
  else
    *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
}

template <typename T, typename CT, bool FP8>
__global__ void __launch_bounds__(256)
    rope_and_cache_kernel(const int64_t* __restrict__ positions, T* __restrict__ query,
                          T* __restrict__ key, const T* __restrict__ value,
                          CT* __restrict__ key_cache, CT* __restrict__ value_cache,
                          const int64_t* __restrict__ slot_mapping,
                          const T* __restrict__ cos_sin_cache, int64_t query_stride,
                          int64_t key_stride, int64_t value_stride, int num_heads,
                          int num_kv_heads, int head_size, int block_size, int num_tokens,
                          int q_ctas, int64_t q_items) {
  constexpr int N = Vec16<T>::N;                   // elements per 16 bytes of T
  constexpr int XE = 16 / sizeof(CT);              // == x of the paged K layout
  const int embed = head_size >> 1;
  const int rchunks = embed / N;                   // rotation chunk pairs per head
  if ((int)blockIdx.x < q_ctas) {
    // ---- query rows ---------------------------------------------------------------------------
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = (int)(item % rchunks);
    const int64_t th = item / rchunks;
    const int head = (int)(th % num_heads);
    const int64_t token = th / num_heads;
    const int64_t pos = item < q_items ? positions[token] : 0;   // step input: ahead of the wait
    pdl_wait();
    pdl_trigger();
    if (item >= q_items) return;
    const T* cache = cos_sin_cache + pos * head_size + chunk * N;
    T* arr = query + token * query_stride + (int64_t)head * head_size + chunk * N;
    Vec16<T> x, y, c, s;
    *reinterpret_cast<uint4*>(&c) = *reinterpret_cast<const uint4*>(cache);
    *reinterpret_cast<uint4*>(&s) = *reinterpret_cast<const uint4*>(cache + embed);
    *reinterpret_cast<uint4*>(&x) = *reinterpret_cast<const uint4*>(arr);
    *reinterpret_cast<uint4*>(&y) = *reinterpret_cast<const uint4*>(arr + embed);
#pragma unroll
    for (int j = 0; j < N; ++j) rot_pair<T>(x.v[j], y.v[j], c.v[j], s.v[j]);
    st_16(arr, *reinterpret_cast<uint4*>(&x));
    st_16(arr + embed, *reinterpret_cast<uint4*>(&y));
    return;
  }
  // ---- key / value rows: one warp per (token, kv head), the first token of a run works ----------
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)((int)blockIdx.x - q_ctas) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t num_rows = (int64_t)num_tokens * num_kv_heads;
  const int64_t token = row < num_rows ? row / num_kv_heads : 0;
  int64_t slot = -1, prev = -2, pos_l = 0;
  int n = 0;
  if (row < num_rows) {
    slot = slot_mapping[token];
    if (token > 0) prev = slot_mapping[token - 1];
    // positions of the run's tokens, one per lane (a run is at most block_size <= 32 tokens):
    // step inputs like the slots, fetched ahead of the wait - the rotation's cos / sin row
    // depends on them
    if (token + lane < num_tokens) pos_l = positions[token + lane];
    const int room = block_size - (int)(slot % block_size);
    const bool cont = lane > 0 && lane < room && token + lane < num_tokens &&
                      slot_mapping[token + lane] == slot + lane;
    const unsigned m = __ballot_sync(0xffffffffu, cont || lane == 0);
    n = __ffs(~m) - 1;
    if (n < 0) n = 32;
    if (n > room) n = room;
  }
  pdl_wait();
  pdl_trigger();
  if (row >= num_rows) return;
  const int head = (int)(row % num_kv_heads);
  const int off = slot >= 0 ? (int)(slot % block_size) : 0;
  const bool in_run = slot >= 0 && prev >= 0 && prev == slot - 1 && off != 0;
  if (in_run) return;                 // the warp of the run's first token handles this token
  if (slot < 0) {
    // padding token: no cache write (cache_kernels.cu:172-175), but its key is still rotated
    // (rotary_embedding has no notion of padding)
    n = 1;
  }
  const int64_t bh = slot >= 0 ? (slot / block_size) * num_kv_heads + head : 0;
  const int chunks = head_size / XE;
  T* kbase = key + token * key_stride + (int64_t)head * head_size;
  const T* vbase = value + token * value_stride + (int64_t)head * head_size;
  // V first: its stores are the long pole
  if (slot >= 0) {
    for (int d = lane; d < head_size; d += 32) {
      CT* vrow = value_cache + (bh * head_size + d) * block_size;
      for (int e0 = (off / XE) * XE; e0 < off + n; e0 += XE) {
        if (e0 >= off && e0 + XE <= off + n) {
          alignas(16) CT buf[XE];
#pragma unroll
          for (int e = 0; e < XE; ++e)
            buf[e] = cache_cvt<T, CT, FP8>(vbase[(int64_t)(e0 - off + e) * value_stride + d]);
          st_16(vrow + e0, *reinterpret_cast<uint4*>(buf));
        } else {
          const int lo = e0 > off ? e0 : off, hi = e0 + XE < off + n ? e0 + XE : off + n;
          for (int e = lo; e < hi; ++e)
            vrow[e] = cache_cvt<T, CT, FP8>(vbase[(int64_t)(e - off) * value_stride + d]);
        }
      }
    }
  }
  // K: (token of the run, chunk pair) over the lanes, consecutive lanes on consecutive tokens
  for (int idx0 = 0; idx0 < rchunks * n; idx0 += 32) {      // warp-uniform trip count (shuffle inside)
    const int idx = idx0 + lane;
    const bool live = idx < rchunks * n;
    const int c = live ? idx / n : 0, i = live ? idx - c * n : 0;
    const int64_t pos = __shfl_sync(0xffffffffu, pos_l, i);
    if (!live) continue;
    const T* cache = cos_sin_cache + pos * head_size + c * N;
    T* arr = kbase + (int64_t)i * key_stride + c * N;
    Vec16<T> x, y, cs, sn;
    *reinterpret_cast<uint4*>(&cs) = *reinterpret_cast<const uint4*>(cache);
    *reinterpret_cast<uint4*>(&sn) = *reinterpret_cast<const uint4*>(cache + embed);
    *reinterpret_cast<uint4*>(&x) = *reinterpret_cast<const uint4*>(arr);
    *reinterpret_cast<uint4*>(&y) = *reinterpret_cast<const uint4*>(arr + embed);
#pragma unroll
    for (int j = 0; j < N; ++j) rot_pair<T>(x.v[j], y.v[j], cs.v[j], sn.v[j]);
    st_16(arr, *reinterpret_cast<uint4*>(&x));
    st_16(arr + embed, *reinterpret_cast<uint4*>(&y));
    if (slot >= 0) {
      alignas(16) CT ox[N], oy[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        ox[j] = cache_cvt<T, CT, FP8>(x.v[j]);
        oy[j] = cache_cvt<T, CT, FP8>(y.v[j]);
      }
      // head elements [c*N, c*N + N) and [embed + c*N, ...): N * sizeof(CT) bytes inside the
      // 16-byte chunk (element / XE) of the token
      const int ex = c * N, ey = embed + c * N;
      CT* kx = key_cache + ((bh * chunks + ex / XE) * block_size + off + i) * XE + ex % XE;
      CT* ky = key_cache + ((bh * chunks + ey / XE) * block_size + off + i) * XE + ey % XE;
      store_bytes<N * (int)sizeof(CT)>(kx, ox);
      store_bytes<N * (int)sizeof(CT)>(ky, oy);
    }
  }
}

template <typename T, typename CT, bool FP8>
static int launch_rope_cache(const int64_t* positions, void* query, void* key, const void* value,
                             void* key_cache, void* value_cache, const int64_t* slot_mapping,
                             const void* cos_sin_cache, int64_t num_tokens, int num_heads,
                             int num_kv_heads, int head_size, int64_t query_stride,
                             int64_t key_stride, int64_t value_stride, int block_size,
                             cudaStream_t stream) {
  constexpr int N = Vec16<T>::N;
  const int rchunks = (head_size / 2) / N;
  const int64_t q_items = num_tokens * num_heads * rchunks;
  const int q_ctas = (int)((q_items + 255) / 256);
  const int64_t rows = num_tokens * num_kv_heads;
  const int kv_ctas = (int)((rows + 7) / 8);
  launch(rope_and_cache_kernel<T, CT, FP8>, dim3((unsigned)(q_ctas + kv_ctas)), dim3(256), 0,
         stream, positions, (T*)query, (T*)key, (const T*)value, (CT*)key_cache,
         (CT*)value_cache, slot_mapping, (const T*)cos_sin_cache, query_stride, key_stride,
         value_stride, num_heads, num_kv_heads, head_size, block_size, (int)num_tokens, q_ctas,
         q_items);
  VLTR_LAUNCH_CHECK("rotary_embedding_and_cache");
  return VLTR_OK;
}

}  // namespace vltr

extern "C" int vltr_rotary_embedding_and_cache(
    const int64_t* positions, void* query, void* key, const void* value, void* key_cache,
    void* value_cache, const int64_t* slot_mapping, const void* cos_sin_cache, int64_t num_tokens,
    int num_heads, int num_kv_heads, int head_size, int rot_dim, int is_neox,
    int64_t query_stride, int64_t key_stride, int64_t value_stride, int block_size, int x,
    int dtype, int kv_fp8, void* stream) {
  using namespace vltr;
  VLTR_CHECK(num_tokens >= 0 && num_heads >= 0 && num_kv_heads >= 0 && head_size > 0,
             VLTR_ERR_INVALID, "rotary_embedding_and_cache: negative size");
  VLTR_CHECK(block_size > 0 && x > 0, VLTR_ERR_INVALID,
             "rotary_embedding_and_cache: block_size %d / x %d", block_size, x);
  if (num_tokens == 0) return VLTR_OK;
  const int es = dtype == VLTR_F32 ? 4 : 2;
  const int n = 16 / es;
  const int xe = kv_fp8 ? 16 : n;
  const bool ok = is_neox && rot_dim == head_size && (head_size / 2) % n == 0 &&
                  head_size % xe == 0 && x == xe && block_size % xe == 0 && block_size <= 32 &&
                  query_stride % n == 0 && key_stride % n == 0 && value_stride % n == 0 &&
                  (((uintptr_t)query | (uintptr_t)key | (uintptr_t)value | (uintptr_t)key_cache |
                    (uintptr_t)value_cache | (uintptr_t)cos_sin_cache) & 15) == 0 &&
                  num_tokens < (1 << 24);
  if (!ok)
    return fail(VLTR_ERR_UNSUPPORTED,
                "rotary_embedding_and_cache: fused form needs NeoX rotation over the whole head "
                "and 16-byte aligned operands (issue rotary_embedding + reshape_and_cache)");
  cudaStream_t s = (cudaStream_t)stream;
  VLTR_DISPATCH_FLOAT(dtype, "rotary_embedding_and_cache", {
    if (kv_fp8)
      return launch_rope_cache<scalar_t, uint8_t, true>(
          positions, query, key, value, key_cache, value_cache, slot_mapping, cos_sin_cache,
          num_tokens, num_heads, num_kv_heads, head_size, query_stride, key_stride, value_stride,
          block_size, s);
    return launch_rope_cache<scalar_t, scalar_t, false>(
        positions, query, key, value, key_cache, value_cache, slot_mapping, cos_sin_cache,
        num_tokens, num_heads, num_kv_heads, head_size, query_stride, key_stride, value_stride,
        block_size, s);
  });
  return VLTR_OK;
}
