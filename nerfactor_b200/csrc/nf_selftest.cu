// Diagnostics (no reference counterpart): throughput of the TMEM read / write path
// (tcgen05.ld / tcgen05.st) on one SM, alone and under a concurrent tcgen05.mma stream.
// The fused MLP kernels' epilogues read every fp32 accumulator tile back through this path, so
// its bytes / clock set their roofline.
#include "nf_common.cuh"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nftc;

#define TC_LD16(r, addr)                                                                        \
  asm volatile(                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                 \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),     \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
        "=r"(r[14]), "=r"(r[15])                                                                \
      : "r"(addr))

// mode bits: 1 = readers run tcgen05.ld.x32, 2 = readers run tcgen05.st.x16, 4 = MMA stream
// (TS mode, 128x128x16, fp16), 8 = loads as .x16 instead of .x32, 16 = MMA stream with N = 256
__global__ void __launch_bounds__(32 * 18, 1)
tmem_bw_kernel(int reader_warps, int iters, int mma_iters, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 2);
  long long* s_t = reinterpret_cast<long long*>(bar + 4);       // [0] readers, [1] mma
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    s_t[0] = 0; s_t[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // zero-ish initial contents so the MMA has defined operands
  if (warp >= 2 && warp < 6) {
    const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0x3c003c00u;
    for (int c = 0; c < 512; c += 16) TC_ST16(tb + c, pk);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long t0 = clock64();
  if (warp == 0) {
    if ((mode & 4) && elect_one()) {
      const int n = (mode & 16) ? 256 : 128;
      const uint32_t idesc = make_idesc(0, n);
      const uint32_t lbo = 128 * 16, sbo = 128;
      for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          tc_mma_ts(tmem_base + 256, tmem_base + 192 + k * 8,
                    make_b_desc(smem_u32(smem) + k * 2 * lbo, lbo, sbo), idesc, 1u);
      }
      tc_commit(bar);
      mbar_wait(bar, 0);
      s_t[1] = clock64() - t0;
    }
  } else if (warp >= 2 && warp < 2 + reader_warps) {
    const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    if (mode & 1) {
      if (mode & 8) {
        for (int it = 0; it < iters; ++it) {
          uint32_t r0[16], r1[16], r2[16], r3[16];
          const uint32_t c = (uint32_t)((it * 64) & 127);
          TC_LD16(r0, tb + c); TC_LD16(r1, tb + c + 16); TC_LD16(r2, tb + c + 32); TC_LD16(r3, tb + c + 48);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
        }
      } else {
        for (int it = 0; it < iters; ++it) {
          uint32_t r0[32], r1[32];
          const uint32_t c = (uint32_t)((it * 64) & 127);
          TC_LD32(r0, tb + c); TC_LD32(r1, tb + c + 32);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= r0[i] ^ r1[i];
        }
      }
    }
    if (mode & 2) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = 0x3c003c00u + (uint32_t)i;
      for (int it = 0; it < iters; ++it) {
        const uint32_t c = (uint32_t)((it * 64) & 127);
        TC_ST16(tb + c, pk); TC_ST16(tb + c + 16, pk); TC_ST16(tb + c + 32, pk); TC_ST16(tb + c + 48, pk);
        tc_wait_st();
      }
    }
    if (acc == 0x12345678u) out[3] = acc;       // keep the loads alive
    const long long dt = clock64() - t0;
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(s_t), (unsigned long long)dt);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = s_t[0]; out[1] = s_t[1]; }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

// out_d[0] = reader cycles (max over reader warps), out_d[1] = MMA-stream cycles.  Every reader
// iteration moves 64 columns x 32 lanes x 4 B = 8 KB per warp; every MMA iteration is 8
// 128 x N x 16 instructions.
extern "C" int nf_selftest_tmem(nf_ctx* ctx, int reader_warps, int iters, int mma_iters, int mode,
                                long long* out_d, void* stream) {
  NF_CHECK_ARG(ctx, out_d && reader_warps >= 0 && reader_warps <= 16 && iters >= 0 && mma_iters >= 0,
               "bad argument");
  const size_t sm = 65536 + 256;
  NF_CUDA(ctx, cudaFuncSetAttribute(tmem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  tmem_bw_kernel<<<1, 32 * 18, sm, (cudaStream_t)stream>>>(reader_warps, iters, mma_iters, mode, out_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
