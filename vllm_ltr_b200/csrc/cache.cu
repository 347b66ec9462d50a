// cache.cu — reshape_and_cache, copy_blocks, swap_blocks, convert_fp8.
//
// Replaces csrc/cache_kernels.cu of the reference: swap_blocks :23-62, copy_blocks :68-152,
// reshape_and_cache :157-276, convert_fp8 :281-339.  All are pure index/copy work and are
// bit-exact with the reference (fp8 goes through the same cuda_fp8.h conversions).
//
// Paged layout (part of the ABI, vllm/attention/ops/paged_attn.py:42-61):
//   key_cache   [num_blocks, num_kv_heads, head_size/x, block_size, x]   x = 16 / sizeof(cache elt)
//   value_cache [num_blocks, num_kv_heads, head_size, block_size]
//
// B200 shape:
//   reshape_and_cache — one warp per (token, head) row flattened over the grid; tokens whose slots
//     are consecutive inside a KV block (every prefill chunk) are written by the warp of the run's
//     first token: K as 16-byte pieces that fill whole lines, V as 16-byte pieces of the
//     block_size-token rows (full sectors).  A lone token (decode) keeps one 16-byte store per K
//     chunk and one element store per V row (the layout strides a token's elements by
//     block_size: one 32-byte sector per element, inherent).
//   copy_blocks — layer pointers and (src,dst) pairs travel as kernel parameters (CUDA >= 12.1
//     allows 32 KB of them): no pointer-array tensors, no blocking H2D copies (the reference does
//     three per call, cache_kernels.cu:129-136), 128-bit copies, grid sized over
//     (slab, layer*2, pair).
//   swap_blocks — pinned host memory is device-mapped under UVA, so a swap of any number of
//     blocks is ONE gather/scatter kernel over PCIe instead of one cudaMemcpyAsync per block
//     (cache_kernels.cu:50-61); pageable memory falls back to memcpy per run of consecutive blocks.
#include "common.cuh"

namespace vltr {

// ------------------------------------------------------------------------------------------------
// reshape_and_cache
// ------------------------------------------------------------------------------------------------
// vector path: one warp per (token, head) row; the warp of the FIRST token of a run does the work.
//   A run = consecutive tokens whose slots are consecutive inside one KV block - what every
//   prefill chunk produces (block_size tokens per block), while decode tokens are runs of one.
//   The paged layouts keep the tokens of a block adjacent (K: 16-byte chunks of consecutive
//   tokens follow each other; V: the block_size tokens of one head element are one contiguous
//   row), so a run turns the scatter into wide stores:
//   K: (chunk, token) pairs over the lanes, consecutive lanes -> consecutive tokens of a chunk:
//      128-bit loads, 128-bit stores that fill whole lines.
//   V: lanes on consecutive head elements; each lane gathers its element from the run's tokens
//      and stores 16 bytes at a time where the run covers an aligned 16-byte group - full 32-byte
//      sectors instead of one read-modify-write sector per element (the single-token case keeps
//      the element stores: inherent to the layout).
template <typename T, typename CT, bool FP8>
__global__ void __launch_bounds__(256)
    reshape_and_cache_vec_kernel(const T* __restrict__ key, const T* __restrict__ value,
                                 CT* __restrict__ key_cache, CT* __restrict__ value_cache,
                                 const int64_t* __restrict__ slot_mapping, int64_t key_stride,
                                 int64_t value_stride, int num_heads, int head_size,
                                 int block_size, int64_t num_rows, int num_tokens) {
  constexpr int XE = 16 / sizeof(CT);              // == x, also tokens per 16 bytes of a V row
  constexpr int LD = XE * sizeof(T) / 16;          // 16-byte loads per K chunk
  constexpr int N = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // the slot mapping is a step input (host-built, copied in before the layer loop): its loads
  // overlap the tail of the kernel that produced key / value
  const int64_t token = row < num_rows ? row / num_heads : 0;
  int64_t slot = -1, prev = -2;
  int n = 0;
  if (row < num_rows) {
    slot = slot_mapping[token];
    if (token > 0) prev = slot_mapping[token - 1];
    // run length: following tokens with consecutive slots, up to the end of the block
    const int room = block_size - (int)(slot % block_size);
    const bool cont = lane > 0 && lane < room && token + lane < num_tokens &&
                      slot_mapping[token + lane] == slot + lane;
    const unsigned m = __ballot_sync(0xffffffffu, cont || lane == 0);
    n = __ffs(~m) - 1;                             // leading ones (lane 0 always set); 32 -> -1
    if (n < 0) n = 32;
    if (n > room) n = room;
  }
  pdl_wait();
  pdl_trigger();
  if (row >= num_rows || slot < 0) return;         // padding token (cache_kernels.cu:172-175)
  const int off = (int)(slot % block_size);
  if (prev >= 0 && prev == slot - 1 && off != 0) return;   // inside a run: its first token's warp writes it
  const int head = (int)(row % num_heads);
  const int64_t blk = slot / block_size;
  const int chunks = head_size / XE;
  const int64_t bh = blk * num_heads + head;
  const T* ksrc = key + token * key_stride + (int64_t)head * head_size;
  const T* vsrc = value + token * value_stride + (int64_t)head * head_size;
  // V first: its stores are the long pole, get them in flight early
  for (int d = lane; d < head_size; d += 32) {
    CT* vrow = value_cache + (bh * head_size + d) * block_size;
    for (int e0 = (off / XE) * XE; e0 < off + n; e0 += XE) {
      if (e0 >= off && e0 + XE <= off + n) {       // the run covers this aligned 16-byte group
        alignas(16) CT buf[XE];
#pragma unroll
        for (int e = 0; e < XE; ++e)
          buf[e] = cache_cvt<T, CT, FP8>(vsrc[(int64_t)(e0 - off + e) * value_stride + d]);
        st_16(vrow + e0, *reinterpret_cast<uint4*>(buf));
      } else {
        const int lo = e0 > off ? e0 : off, hi = e0 + XE < off + n ? e0 + XE : off + n;
        for (int e = lo; e < hi; ++e)
          vrow[e] = cache_cvt<T, CT, FP8>(vsrc[(int64_t)(e - off) * value_stride + d]);
      }
    }
  }
  for (int idx = lane; idx < chunks * n; idx += 32) {
    const int chunk = idx / n, i = idx - chunk * n;
    alignas(16) T kbuf[XE];
#pragma unroll
    for (int l = 0; l < LD; ++l)
      *reinterpret_cast<uint4*>(&kbuf[l * N]) =
          ld_nc_16(ksrc + (int64_t)i * key_stride + chunk * XE + l * N);
    alignas(16) CT kout[XE];
#pragma unroll
    for (int e = 0; e < XE; ++e) kout[e] = cache_cvt<T, CT, FP8>(kbuf[e]);
    CT* kdst = key_cache + ((bh * chunks + chunk) * block_size + off + i) * XE;
    st_16(kdst, *reinterpret_cast<uint4*>(kout));
  }
}

// scalar path: one element per thread, any x / alignment
template <typename T, typename CT, bool FP8>
__global__ void __launch_bounds__(256)
    reshape_and_cache_scalar_kernel(const T* __restrict__ key, const T* __restrict__ value,
                                    CT* __restrict__ key_cache, CT* __restrict__ value_cache,
                                    const int64_t* __restrict__ slot_mapping, int64_t key_stride,
                                    int64_t value_stride, int num_heads, int head_size,
                                    int block_size, int x, int64_t num_items) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= num_items) return;
  const int n = num_heads * head_size;
  const int i = (int)(item % n);
  const int64_t token = item / n;
  const int64_t slot = slot_mapping[token];
  if (slot < 0) return;
  const int64_t blk = slot / block_size;
  const int off = (int)(slot % block_size);
  const int head = i / head_size, ho = i % head_size;
  const int x_idx = ho / x, x_off = ho % x;
  const int64_t bh = blk * num_heads + head;
  const int64_t kt = ((bh * (head_size / x) + x_idx) * block_size + off) * x + x_off;
  const int64_t vt = (bh * head_size + ho) * block_size + off;
  key_cache[kt] = cache_cvt<T, CT, FP8>(key[token * key_stride + i]);
  value_cache[vt] = cache_cvt<T, CT, FP8>(value[token * value_stride + i]);
}

template <typename T, typename CT, bool FP8>
int launch_reshape(const void* key, const void* value, void* key_cache, void* value_cache,
                   const int64_t* slot_mapping, int num_tokens, int num_heads, int head_size,
                   int block_size, int x, int64_t key_stride, int64_t value_stride,
                   cudaStream_t stream) {
  constexpr int XE = 16 / sizeof(CT);
  constexpr int N = 16 / sizeof(T);
  if (num_tokens == 0 || num_heads == 0 || head_size == 0) return VLTR_OK;
  const bool vec_ok = x == XE && head_size % XE == 0 && key_stride % N == 0 &&
                      value_stride % N == 0 && (uintptr_t)key % 16 == 0 &&
                      (uintptr_t)value % 16 == 0 && (uintptr_t)key_cache % 16 == 0 &&
                      (uintptr_t)value_cache % 16 == 0 && block_size % XE == 0 && block_size <= 32;
  if (vec_ok) {
    const int64_t rows = (int64_t)num_tokens * num_heads;
    launch(reshape_and_cache_vec_kernel<T, CT, FP8>, dim3((unsigned)((rows + 7) / 8)), dim3(256),
           0, stream, (const T*)key, (const T*)value, (CT*)key_cache, (CT*)value_cache,
           slot_mapping, key_stride, value_stride, num_heads, head_size, block_size, rows,
           num_tokens);
  } else {
    const int64_t items = (int64_t)num_tokens * num_heads * head_size;
    reshape_and_cache_scalar_kernel<T, CT, FP8>
        <<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(
            (const T*)key, (const T*)value, (CT*)key_cache, (CT*)value_cache, slot_mapping,
            key_stride, value_stride, num_heads, head_size, block_size, x, items);
  }
  VLTR_LAUNCH_CHECK("reshape_and_cache");
  return VLTR_OK;
}

// ------------------------------------------------------------------------------------------------
// block copies (copy_blocks / swap_blocks)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxBufs = 256;    // 2 * num_layers pointers per launch
constexpr int kMaxPairs = 512;   // (src,dst) pairs per launch

struct CopyParams {
  const char* src[kMaxBufs];
  char* dst[kMaxBufs];
  int64_t pair_src[kMaxPairs];
  int64_t pair_dst[kMaxPairs];
  int64_t block_bytes;
};

// grid (slabs, bufs, pairs); each CTA copies a 16 KiB slab of one block of one buffer
constexpr int kSlabBytes = 16384;
__global__ void __launch_bounds__(256) block_copy_kernel(const __grid_constant__ CopyParams p) {
  const char* src = p.src[blockIdx.y] + p.pair_src[blockIdx.z] * p.block_bytes;
  char* dst = p.dst[blockIdx.y] + p.pair_dst[blockIdx.z] * p.block_bytes;
  const int64_t begin = (int64_t)blockIdx.x * kSlabBytes;
  const int64_t end = min(begin + (int64_t)kSlabBytes, p.block_bytes);
  if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)p.block_bytes) & 15) == 0) {
    uint4 v[4];
    const int64_t o = begin + (int64_t)threadIdx.x * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (o + k * 4096 < end) v[k] = *reinterpret_cast<const uint4*>(src + o + k * 4096);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (o + k * 4096 < end) *reinterpret_cast<uint4*>(dst + o + k * 4096) = v[k];
  } else {
    for (int64_t o = begin + threadIdx.x; o < end; o += blockDim.x) dst[o] = src[o];
  }
}

static int run_block_copies(const char* const* src, char* const* dst, int nbufs,
                            const int64_t* mapping, int num_pairs, int64_t block_bytes,
                            cudaStream_t stream, const char* name) {
  if (nbufs == 0 || num_pairs == 0 || block_bytes == 0) return VLTR_OK;
  const unsigned slabs = (unsigned)((block_bytes + kSlabBytes - 1) / kSlabBytes);
  for (int b0 = 0; b0 < nbufs; b0 += kMaxBufs) {
    const int nb = nbufs - b0 < kMaxBufs ? nbufs - b0 : kMaxBufs;
    for (int p0 = 0; p0 < num_pairs; p0 += kMaxPairs) {
      const int np = num_pairs - p0 < kMaxPairs ? num_pairs - p0 : kMaxPairs;
      CopyParams p;
      for (int i = 0; i < nb; ++i) {
        p.src[i] = src[b0 + i];
        p.dst[i] = dst[b0 + i];
      }
      for (int i = 0; i < np; ++i) {
        p.pair_src[i] = mapping[2 * (p0 + i)];
        p.pair_dst[i] = mapping[2 * (p0 + i) + 1];
      }
      p.block_bytes = block_bytes;
      block_copy_kernel<<<dim3(slabs, nb, np), 256, 0, stream>>>(p);
      VLTR_LAUNCH_CHECK(name);
    }
  }
  return VLTR_OK;
}

// ------------------------------------------------------------------------------------------------
// convert_fp8
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    to_fp8_kernel(const T* __restrict__ src, uint8_t* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = to_fp8<T>(src[i]);
}
template <typename T>
__global__ void __launch_bounds__(256)
    from_fp8_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = from_fp8<T>(src[i]);
}

}  // namespace vltr

extern "C" int vltr_reshape_and_cache(const void* key, const void* value, void* key_cache,
                                      void* value_cache, const int64_t* slot_mapping,
                                      int num_tokens, int num_heads, int head_size,
                                      int block_size, int x, int64_t key_stride,
                                      int64_t value_stride, int dtype, int kv_fp8, float kv_scale,
                                      void* stream) {
  using namespace vltr;
  (void)kv_scale;  // unscaled e5m2: ignored on CUDA, as in the reference
  VLTR_CHECK(num_tokens >= 0 && num_heads >= 0 && head_size >= 0, VLTR_ERR_INVALID,
             "reshape_and_cache: negative size");
  VLTR_CHECK(block_size > 0 && x > 0 && head_size % x == 0, VLTR_ERR_INVALID,
             "reshape_and_cache: block_size %d / x %d invalid for head_size %d", block_size, x,
             head_size);
  cudaStream_t s = (cudaStream_t)stream;
  VLTR_DISPATCH_FLOAT(dtype, "reshape_and_cache", {
    if (kv_fp8)
      return launch_reshape<scalar_t, uint8_t, true>(key, value, key_cache, value_cache,
                                                     slot_mapping, num_tokens, num_heads,
                                                     head_size, block_size, x, key_stride,
                                                     value_stride, s);
    return launch_reshape<scalar_t, scalar_t, false>(key, value, key_cache, value_cache,
                                                     slot_mapping, num_tokens, num_heads,
                                                     head_size, block_size, x, key_stride,
                                                     value_stride, s);
  });
  return VLTR_OK;
}

extern "C" int vltr_copy_blocks(void* const* key_cache_ptrs, void* const* value_cache_ptrs,
                                int num_layers, const int64_t* block_mapping, int num_pairs,
                                int64_t block_bytes, void* stream) {
  using namespace vltr;
  VLTR_CHECK(num_layers >= 0 && num_pairs >= 0 && block_bytes >= 0, VLTR_ERR_INVALID,
             "copy_blocks: negative size");
  if (num_layers == 0 || num_pairs == 0) return VLTR_OK;
  // one buffer list: K of every layer then V of every layer; src and dst are the same cache
  const int nbufs = 2 * num_layers;
  const char** src = (const char**)alloca(sizeof(char*) * nbufs);
  char** dst = (char**)alloca(sizeof(char*) * nbufs);
  for (int l = 0; l < num_layers; ++l) {
    src[l] = dst[l] = (char*)key_cache_ptrs[l];
    src[num_layers + l] = dst[num_layers + l] = (char*)value_cache_ptrs[l];
  }
  return run_block_copies(src, dst, nbufs, block_mapping, num_pairs, block_bytes,
                          (cudaStream_t)stream, "copy_blocks");
}

extern "C" int vltr_swap_blocks(const void* src, void* dst, int kind,
                                const int64_t* block_mapping, int num_pairs, int64_t block_bytes,
                                void* stream) {
  using namespace vltr;
  VLTR_CHECK(kind >= 0 && kind <= 2, VLTR_ERR_INVALID, "swap_blocks: Invalid device combination");
  VLTR_CHECK(num_pairs >= 0 && block_bytes >= 0, VLTR_ERR_INVALID, "swap_blocks: negative size");
  if (num_pairs == 0 || block_bytes == 0) return VLTR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  bool mapped = true;
  if (kind != 0) {
    // the host side must be pinned (cudaHostAlloc / cudaHostRegister) to be read by a kernel
    const void* host = kind == 1 ? dst : src;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, host);
    if (e != cudaSuccess) {
      cudaGetLastError();
      mapped = false;
    } else {
      mapped = attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
      if (mapped) {
        if (kind == 1)
          dst = (char*)attr.devicePointer;
        else
          src = (const char*)attr.devicePointer;
      }
    }
  }
  if (mapped) {
    const char* sp = (const char*)src;
    char* dp = (char*)dst;
    return run_block_copies(&sp, &dp, 1, block_mapping, num_pairs, block_bytes, s, "swap_blocks");
  }
  // pageable host memory: cudaMemcpyAsync per run of consecutive (src,dst) blocks
  const cudaMemcpyKind mk = kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
  int i = 0;
  while (i < num_pairs) {
    int j = i + 1;
    while (j < num_pairs && block_mapping[2 * j] == block_mapping[2 * (j - 1)] + 1 &&
           block_mapping[2 * j + 1] == block_mapping[2 * (j - 1) + 1] + 1)
      ++j;
    VLTR_CUDA(cudaMemcpyAsync((char*)dst + block_mapping[2 * i + 1] * block_bytes,
                              (const char*)src + block_mapping[2 * i] * block_bytes,
                              (size_t)(j - i) * block_bytes, mk, s));
    i = j;
  }
  return VLTR_OK;
}

extern "C" int vltr_swap_blocks_layers(const void* const* src_ptrs, void* const* dst_ptrs,
                                       int num_bufs, int kind, const int64_t* block_mapping,
                                       int num_pairs, int64_t block_bytes, void* stream) {
  using namespace vltr;
  VLTR_CHECK(kind >= 0 && kind <= 2, VLTR_ERR_INVALID, "swap_blocks: Invalid device combination");
  VLTR_CHECK(num_bufs >= 0 && num_pairs >= 0 && block_bytes >= 0, VLTR_ERR_INVALID,
             "swap_blocks: negative size");
  if (num_bufs == 0 || num_pairs == 0 || block_bytes == 0) return VLTR_OK;
  const char** src = (const char**)alloca(sizeof(char*) * num_bufs);
  char** dst = (char**)alloca(sizeof(char*) * num_bufs);
  bool mapped = true;
  for (int i = 0; i < num_bufs && mapped; ++i) {
    src[i] = (const char*)src_ptrs[i];
    dst[i] = (char*)dst_ptrs[i];
    if (kind != 0) {
      // the host side must be pinned (device-mapped under UVA) to be read / written by a kernel
      cudaPointerAttributes attr;
      const void* host = kind == 1 ? (const void*)dst[i] : (const void*)src[i];
      if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
        cudaGetLastError();
        mapped = false;
      } else if (attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) {
        mapped = false;
      } else if (kind == 1) {
        dst[i] = (char*)attr.devicePointer;
      } else {
        src[i] = (const char*)attr.devicePointer;
      }
    }
  }
  if (mapped)   // every block of every layer in one gather / scatter launch
    return run_block_copies(src, dst, num_bufs, block_mapping, num_pairs, block_bytes,
                            (cudaStream_t)stream, "swap_blocks_layers");
  for (int i = 0; i < num_bufs; ++i) {   // pageable host memory: buffer by buffer, memcpy runs
    const int rc = vltr_swap_blocks(src_ptrs[i], dst_ptrs[i], kind, block_mapping, num_pairs,
                                    block_bytes, stream);
    if (rc != VLTR_OK) return rc;
  }
  return VLTR_OK;
}

extern "C" int vltr_convert_fp8(const void* src, void* dst, int64_t numel, int src_dtype,
                                int dst_dtype, void* stream) {
  using namespace vltr;
  VLTR_CHECK(numel >= 0, VLTR_ERR_INVALID, "convert_fp8: negative size");
  VLTR_CHECK((src_dtype == VLTR_U8) != (dst_dtype == VLTR_U8), VLTR_ERR_INVALID,
             "convert_fp8: exactly one of src/dst must be the fp8 (uint8) cache");
  if (numel == 0) return VLTR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((numel + 255) / 256);
  if (dst_dtype == VLTR_U8) {
    VLTR_DISPATCH_FLOAT(src_dtype, "convert_fp8", {
      to_fp8_kernel<scalar_t><<<grid, 256, 0, s>>>((const scalar_t*)src, (uint8_t*)dst, numel);
    });
  } else {
    VLTR_DISPATCH_FLOAT(dst_dtype, "convert_fp8", {
      from_fp8_kernel<scalar_t><<<grid, 256, 0, s>>>((const uint8_t*)src, (scalar_t*)dst, numel);
    });
  }
  VLTR_LAUNCH_CHECK("convert_fp8");
  return VLTR_OK;
}
