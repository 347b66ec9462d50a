// pos_encoding.cu — rotary_embedding / batched_rotary_embedding (in place on query and key).
//
// Replaces csrc/pos_encoding_kernels.cu:10-120 (kernels) and :124-226 (launchers).
//
// Reference numerics (kept, so f16/bf16 are bit-identical): the arithmetic is "in T" —
//   arr[x] = x*cos - y*sin ; arr[y] = y*cos + x*sin     with every * and +/- rounded to T
// (pos_encoding_kernels.cu:34-38; c10::Half/BFloat16 operators round after each op).
//
// B200 shape: tiny streaming op, T*(Hq+Hkv)*D*s*2 bytes.  NeoX style (Llama) is vectorised:
// one thread rotates 16 bytes of the first half against 16 bytes of the second half, and the
// (token, head, chunk) space is flattened over the grid so 256 decode tokens give >100k
// independent 64-byte work items instead of 256 CTAs of strided 2-byte accesses.
#include "common.cuh"

namespace vltr {

// NeoX, vectorised.  item = ((token * (Hq+Hkv)) + head) * chunks + chunk
template <typename T>
__global__ void __launch_bounds__(256)
    rope_neox_vec_kernel(const int64_t* __restrict__ positions, T* __restrict__ query,
                         T* __restrict__ key, const T* __restrict__ cos_sin_cache,
                         const int64_t* __restrict__ offsets, int rot_dim, int64_t query_stride,
                         int64_t key_stride, int num_heads, int num_kv_heads, int head_size,
                         int64_t num_items) {
  constexpr int N = Vec16<T>::N;
  pdl_wait();
  pdl_trigger();
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= num_items) return;
  const int embed = rot_dim >> 1;
  const int chunks = embed / N;
  const int heads = num_heads + num_kv_heads;
  const int chunk = (int)(item % chunks);
  const int64_t th = item / chunks;
  const int head = (int)(th % heads);
  const int64_t token = th / heads;
  int64_t pos = positions[token];
  if (offsets != nullptr) pos += offsets[token];
  const T* cache = cos_sin_cache + pos * rot_dim + chunk * N;
  T* arr = head < num_heads ? query + token * query_stride + (int64_t)head * head_size
                            : key + token * key_stride + (int64_t)(head - num_heads) * head_size;
  arr += chunk * N;
  Vec16<T> x, y, c, s;
  *reinterpret_cast<uint4*>(&c) = *reinterpret_cast<const uint4*>(cache);
  *reinterpret_cast<uint4*>(&s) = *reinterpret_cast<const uint4*>(cache + embed);
  *reinterpret_cast<uint4*>(&x) = *reinterpret_cast<const uint4*>(arr);
  *reinterpret_cast<uint4*>(&y) = *reinterpret_cast<const uint4*>(arr + embed);
#pragma unroll
  for (int j = 0; j < N; ++j) rot_pair<T>(x.v[j], y.v[j], c.v[j], s.v[j]);
  st_16(arr, *reinterpret_cast<uint4*>(&x));
  st_16(arr + embed, *reinterpret_cast<uint4*>(&y));
}

// generic: one (token, head, rot_offset) per thread; NeoX or GPT-J
template <typename T, bool IS_NEOX>
__global__ void __launch_bounds__(256)
    rope_scalar_kernel(const int64_t* __restrict__ positions, T* __restrict__ query,
                       T* __restrict__ key, const T* __restrict__ cos_sin_cache,
                       const int64_t* __restrict__ offsets, int rot_dim, int64_t query_stride,
                       int64_t key_stride, int num_heads, int num_kv_heads, int head_size,
                       int64_t num_items) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= num_items) return;
  const int embed = rot_dim >> 1;
  const int heads = num_heads + num_kv_heads;
  const int rot_offset = (int)(item % embed);
  const int64_t th = item / embed;
  const int head = (int)(th % heads);
  const int64_t token = th / heads;
  int64_t pos = positions[token];
  if (offsets != nullptr) pos += offsets[token];
  const T* cache = cos_sin_cache + pos * rot_dim;
  T* arr = head < num_heads ? query + token * query_stride + (int64_t)head * head_size
                            : key + token * key_stride + (int64_t)(head - num_heads) * head_size;
  const int xi = IS_NEOX ? rot_offset : 2 * rot_offset;
  const int yi = IS_NEOX ? embed + rot_offset : 2 * rot_offset + 1;
  T x = arr[xi], y = arr[yi];
  rot_pair<T>(x, y, cache[rot_offset], cache[embed + rot_offset]);
  arr[xi] = x;
  arr[yi] = y;
}

template <typename T>
int launch_rope(const int64_t* positions, void* query, void* key, int64_t num_tokens,
                int num_heads, int num_kv_heads, int head_size, int rot_dim, int64_t query_stride,
                int64_t key_stride, const void* cache, const int64_t* offsets, int is_neox,
                cudaStream_t stream) {
  constexpr int N = Vec16<T>::N;
  const int embed = rot_dim / 2;
  const int heads = num_heads + num_kv_heads;
  if (num_tokens == 0 || heads == 0 || embed == 0) return VLTR_OK;
  const bool vec_ok = is_neox && embed % N == 0 && head_size % N == 0 && query_stride % N == 0 &&
                      key_stride % N == 0 && rot_dim % N == 0 && (uintptr_t)query % 16 == 0 &&
                      (uintptr_t)key % 16 == 0 && (uintptr_t)cache % 16 == 0;
  if (vec_ok) {
    const int64_t items = num_tokens * heads * (embed / N);
    launch(rope_neox_vec_kernel<T>, dim3((unsigned)((items + 255) / 256)), dim3(256), 0, stream,
           positions, (T*)query, (T*)key, (const T*)cache, offsets, rot_dim, query_stride,
           key_stride, num_heads, num_kv_heads, head_size, items);
  } else {
    const int64_t items = num_tokens * heads * embed;
    const unsigned grid = (unsigned)((items + 255) / 256);
    if (is_neox)
      rope_scalar_kernel<T, true><<<grid, 256, 0, stream>>>(
          positions, (T*)query, (T*)key, (const T*)cache, offsets, rot_dim, query_stride,
          key_stride, num_heads, num_kv_heads, head_size, items);
    else
      rope_scalar_kernel<T, false><<<grid, 256, 0, stream>>>(
          positions, (T*)query, (T*)key, (const T*)cache, offsets, rot_dim, query_stride,
          key_stride, num_heads, num_kv_heads, head_size, items);
  }
  VLTR_LAUNCH_CHECK("rotary_embedding");
  return VLTR_OK;
}

}  // namespace vltr

extern "C" int vltr_rotary_embedding(const int64_t* positions, void* query, void* key,
                                     int64_t num_tokens, int num_heads, int num_kv_heads,
                                     int head_size, int rot_dim, int64_t query_stride,
                                     int64_t key_stride, const void* cos_sin_cache,
                                     const int64_t* cos_sin_cache_offsets, int is_neox, int dtype,
                                     void* stream) {
  using namespace vltr;
  VLTR_CHECK(rot_dim >= 0 && rot_dim <= head_size && rot_dim % 2 == 0, VLTR_ERR_INVALID,
             "rotary_embedding: rot_dim %d invalid for head_size %d", rot_dim, head_size);
  VLTR_DISPATCH_FLOAT(dtype, "rotary_embedding", {
    return launch_rope<scalar_t>(positions, query, key, num_tokens, num_heads, num_kv_heads,
                                 head_size, rot_dim, query_stride, key_stride, cos_sin_cache,
                                 cos_sin_cache_offsets, is_neox, (cudaStream_t)stream);
  });
  return VLTR_OK;
}
