// Micro-benchmark: what HBM bandwidth can a B200 kernel reach for (a) a pure read stream through per-warp
// TMA bulk-copy rings (the st_gcn kernel's access pattern, no math), (b) plain LDG.128 reads, (c) plain
// STG.128 writes (the graph kernel's pattern), (d) a device-to-device copy?  These are the practical ceilings
// the st_gcn / graph kernels are compared with (profiles/r2*_hbm_stream.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_stream hbm_stream.cu && ./hbm_stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// every warp streams chunks w, w + W, w + 2W, ... (W = all warps of the grid) of `chunk` bytes through its own ring
template <int STAGES>
__global__ void tma_ring_read(const float* __restrict__ src, size_t n_chunks, int chunk_bytes, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem) + w * STAGES;
  float* ring = reinterpret_cast<float*>(smem + 1024) + (size_t)w * STAGES * (chunk_bytes / 4);
  if (lane == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t W = (size_t)gridDim.x * warps, first = (size_t)blockIdx.x * warps + w;
  size_t nextp = first;
  auto produce = [&](int st) {
    if (nextp < n_chunks && lane == 0) {
      mbar_expect(&bar[st], chunk_bytes);
      bulk_g2s(ring + (size_t)st * (chunk_bytes / 4), reinterpret_cast<const char*>(src) + nextp * chunk_bytes, chunk_bytes, &bar[st]);
    }
    nextp += W;
  };
  for (int i = 0; i < STAGES; ++i) produce(i);
  float acc = 0.f;
  uint32_t k = 0;
  for (size_t c = first; c < n_chunks; c += W, ++k) {
    const int st = k % STAGES;
    mbar_wait(&bar[st], (k / STAGES) & 1);
    acc += ring[(size_t)st * (chunk_bytes / 4) + lane];   // touch the stage
    __syncwarp();
    produce(st);
  }
  if (acc == 123.456f) out[0] = acc;
}

__global__ void ldg_read(const float4* __restrict__ src, size_t n4, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    acc += a.x + b.y + c.z + d.w;
  }
  for (; i < n4; i += stride) acc += __ldcs(src + i).x;
  if (acc == 123.456f) out[0] = acc;
}

__global__ void stg_write(float4* __restrict__ dst, size_t n4, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 q = make_float4(v, v, v, v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = q;
}

// The graph kernel's store pattern: a CTA of `warps` warps takes 64 KB items (grid-stride); PER_WARP = true: every
// warp streams its own contiguous 64 KB / warps slice, 512 B per instruction (what graph_build_fast_kernel does);
// false: the whole CTA writes the item front to back, warps interleaved at 512 B.
template <bool PER_WARP>
__global__ void item_write(float4* __restrict__ dst, size_t n_items, float v) {
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 q = make_float4(v, v, v, v);
  const int item4 = 65536 / 16;            // float4 per item
  const int slice4 = item4 / warps;
  for (size_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    float4* base = dst + it * item4;
    if (PER_WARP) {
      float4* p = base + (size_t)w * slice4 + lane;
      for (int i = 0; i < slice4; i += 32) p[i] = q;
    } else {
      float4* p = base + threadIdx.x;
      for (int i = 0; i < item4; i += blockDim.x) p[i] = q;
    }
  }
}

// The same items written with TMA bulk stores: every warp fills its 16 KB (64 KB / warps) slice in shared memory
// (STS.128), fences the async proxy and issues `parts` cp.async.bulk shared -> global copies.
__global__ void item_write_bulk(float* __restrict__ dst, size_t n_items, float v, int parts) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice_bytes = 65536 / warps;
  float4* tile = reinterpret_cast<float4*>(smem + (size_t)w * slice_bytes);
  const float4 q = make_float4(v, v, v, v);
  const int part_bytes = slice_bytes / parts;
  for (size_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    char* g = reinterpret_cast<char*>(dst) + it * 65536 + (size_t)w * slice_bytes;
    for (int pI = 0; pI < parts; ++pI) {
      // the previous bulk store that read this part must be done reading before it is overwritten
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      float4* t4 = tile + pI * (part_bytes / 16);
      for (int i = lane; i < part_bytes / 16; i += 32) t4[i] = q;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + pI * part_bytes),
                     "r"(smem_u32(t4)), "r"(part_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = src[i];
}

template <class F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t bytes = size_t(16) << 30;   // 16 GiB: far beyond the 126 MB L2
  float *src, *dst, *out;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes)); CK(cudaMalloc(&out, 16));
  CK(cudaMemset(src, 0, bytes)); CK(cudaMemset(dst, 0, bytes));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d, buffer %.1f GB\n", sms, bytes / 1e9);
  {
    float ms = time_ms([&] { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D           : %7.1f GB/s (read + write bytes)\n", 2.0 * bytes / ms / 1e6);
    ms = time_ms([&] { copy_kernel<<<sms * 8, 512>>>((const float4*)src, (float4*)dst, bytes / 16); });
    printf("copy kernel LDG/STG.128  : %7.1f GB/s (read + write bytes)\n", 2.0 * bytes / ms / 1e6);
    ms = time_ms([&] { ldg_read<<<sms * 8, 512>>>((const float4*)src, bytes / 16, out); });
    printf("read  LDG.128 (4 in flight): %7.1f GB/s\n", bytes / ms / 1e6);
    ms = time_ms([&] { stg_write<<<sms * 8, 512>>>((float4*)dst, bytes / 16, 1.f); });
    printf("write STG.128            : %7.1f GB/s\n", bytes / ms / 1e6);
    for (int warps : {4, 8})
      for (int per_sm : {3, 6}) {
        ms = time_ms([&] { item_write<true><<<sms * per_sm, warps * 32>>>((float4*)dst, bytes / 65536, 1.f); });
        printf("write 64 KB items, per-warp slices, %d warps x %d CTAs/SM : %7.1f GB/s\n", warps, per_sm, bytes / ms / 1e6);
        ms = time_ms([&] { item_write<false><<<sms * per_sm, warps * 32>>>((float4*)dst, bytes / 65536, 1.f); });
        printf("write 64 KB items, CTA-contiguous,  %d warps x %d CTAs/SM : %7.1f GB/s\n", warps, per_sm, bytes / ms / 1e6);
      }
    cudaFuncSetAttribute(item_write_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int parts : {1, 2, 4})
      for (int per_sm : {2, 3}) {
        ms = time_ms([&] { item_write_bulk<<<sms * per_sm, 128, 65536>>>(dst, bytes / 65536, 1.f, parts); });
        cudaError_t e = cudaGetLastError();
        printf("write 64 KB items, TMA bulk stores, 4 warps x %d CTAs/SM, %d parts per 16 KB slice : %7.1f GB/s %s\n", per_sm, parts,
               bytes / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    ms = time_ms([&] { cudaMemsetAsync(dst, 0, bytes); });
    printf("cudaMemset               : %7.1f GB/s\n", bytes / ms / 1e6);
  }
  const int chunks[] = {2048, 4096, 4608, 8192, 16384};
  for (int cps = 1; cps <= 2; ++cps)
    for (int warps : {8})
      for (int ci = 0; ci < 5; ++ci) {
        const int chunk = chunks[ci];
        auto run = [&](auto stages_tag) {
          constexpr int ST = decltype(stages_tag)::value;
          const size_t smem = 1024 + (size_t)warps * ST * chunk;
          if (smem * cps > 225 * 1024) return;
          auto k = tma_ring_read<ST>;
          cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          const size_t n_chunks = bytes / chunk;
          float ms = time_ms([&] { k<<<sms * cps, warps * 32, smem>>>(src, n_chunks, chunk, out); }, 3);
          cudaError_t e = cudaGetLastError();
          printf("TMA ring read  CTAs/SM %d  warps %2d  stages %d  chunk %5d B  in flight/SM %6.1f KB : %7.1f GB/s %s\n", cps, warps, ST, chunk,
                 cps * warps * ST * chunk / 1024.0, bytes / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
        run(std::integral_constant<int, 2>{});
        run(std::integral_constant<int, 4>{});
      }
  return 0;
}
