// Bring-up microbenchmark (tests / DESIGN.md measurements only): sustained tcgen05.mma issue rate for
// kind::f16 M=128, N in {64,128,256}, K=16, both operands in shared memory (SWIZZLE_128B, K-major or MN-major).
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

namespace srk {
using namespace tc;

__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int n, int iters, int a_row_off, int mn_major, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A: 256 rows x 128 B, B: 256 rows x 128 B (contents irrelevant: zero)
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, n, mn_major, mn_major);
    const uint64_t hi = make_smem_desc(0, mn_major ? 0 : 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo = (uint32_t)(make_smem_desc(0, mn_major ? 0 : 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t a_lo = lo + ((base + a_row_off * 128) >> 4), b_lo = lo + ((base + 32 * 1024) >> 4);
    const uint32_t kstep = mn_major ? (16 * 128 / 16) : 2;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem, hi | (a_lo + ks * kstep), hi | (b_lo + ks * kstep), idesc, 1);
      }
      umma_commit(smem_u32(&bar));
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr, 0);
    if (elect_one()) {
      t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// out_host[0] = issue cycles per MMA, out_host[1] = completion cycles per MMA
int probe_mma_rate(int n, int a_row_off, int mn_major, float* out_host) {
  long long* d = nullptr;
  if (cudaMalloc(&d, 2 * sizeof(long long)) != cudaSuccess) SRK_FAIL("probe: cudaMalloc failed");
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const int iters = 2000;
  mma_rate_kernel<<<148, 128, 66 * 1024>>>(n, iters, a_row_off, mn_major, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) SRK_FAIL("probe: %s", cudaGetErrorString(e));
  out_host[0] = (float)h[0] / (4.f * iters);
  out_host[1] = (float)h[1] / (4.f * iters);
  return 0;
}

// TMEM read rate: `nwarps` warps (4 or 8; warp w reads lane group w & 3) loop tcgen05.ld.32x32b.x32 (4 KB per
// instruction); `batch` loads are issued back to back before each tcgen05.wait::ld.
__global__ void __launch_bounds__(256, 1) ldtm_rate_kernel(int iters, int batch, long long* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32], acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int b = 0; b < batch; ++b) {
      tmem_ld_32x32(tmem + (((warp >> 2) * 8 + b) & 15) * 32, v);
      if (b == batch - 1) tmem_ld_wait();
    }
    acc += v[0] ^ v[31];
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[2] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_slot, 512); }
}

// out_host[0] = cycles per 4 KB tcgen05.ld per warp, out_host[1] = TMEM bytes read per cycle per SM
int probe_ldtm_rate(int nwarps, int batch, float* out_host) {
  long long* d = nullptr;
  if (cudaMalloc(&d, 4 * sizeof(long long)) != cudaSuccess) SRK_FAIL("probe: cudaMalloc failed");
  const int iters = 2000;
  ldtm_rate_kernel<<<148, nwarps * 32>>>(iters, batch, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) SRK_FAIL("probe: %s", cudaGetErrorString(e));
  out_host[0] = (float)h[0] / ((float)iters * batch);
  out_host[1] = 4096.f * nwarps * iters * batch / (float)h[0];
  return 0;
}
}  // namespace srk
