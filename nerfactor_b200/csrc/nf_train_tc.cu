// tcgen05 Dense layers for the training step (config 4: forward + backward, 16-bit operands,
// fp32 accumulation, fp32 master weights / activations / gradients in HBM).
//
// One Keras Dense of mlp.Network (nerfactor/networks/mlp.py:34, 39-50) and what tape.gradient
// (nerfactor/trainvali.py:278-285) computes for it:
//     y  = act([x1 | x2] W + b)                      rowgemm_tc_kernel<.., A16 = 0>
//     dz = dy * act'(y), db += colsum(dz)            act_bwd_colsum_kernel (writes dz as 16 bit)
//     dx1 | dx2 = dz W^T                             rowgemm_tc_kernel<.., A16 = 1>
//     dW += [x1 | x2]^T dz                           wgrad_tc_kernel (+ reduce_partials_kernel)
// All four are HBM-bound streaming kernels over the [rows, features] activations (rows =
// (ray, light) pairs, 0.5 M per step); the tensor pipe is far from saturated, so the design
// goal is simply to touch every activation byte once per kernel:
//  * rowgemm: persistent CTA per SM; the (tiny) weight operand is resident in shared memory as
//    a swizzle-free K-major image; 4 loader warps convert 128-row fp32 tiles to 16-bit K-major
//    A images (double buffered), one thread issues tcgen05.mma (SS mode) into a double-buffered
//    TMEM accumulator, 4 epilogue warps apply bias / activation and store fp32.
//  * wgrad: both operands are the SAME row-major tiles read as MN-major operands (the
//    contraction runs over the rows), accumulated over all of a CTA's tiles in TMEM and
//    written once as a per-CTA partial; a deterministic second pass sums the partials.
#include "nf_common.cuh"
#include "nf_tc_ptx.cuh"
#include <algorithm>
#include <type_traits>
#include <cstdlib>

namespace {
using namespace nftc;

#define TC_LD16(r, addr)                                                                         \
  asm volatile(                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                  \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"           \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),      \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),  \
        "=r"(r[14]), "=r"(r[15])                                                                 \
      : "r"(addr)                                                                                \
      : "memory")

__device__ __forceinline__ float act_grad_tc(int act, float y) {
  switch (act) {
    case NF_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case NF_ACT_SIGMOID: return y * (1.f - y);
    case NF_ACT_SOFTPLUS: return 1.f - expf(-y);
    default: return 1.f;
  }
}

// instruction descriptor with operand-major bits (cute UMMA::InstrDescriptor: a_major bit 15,
// b_major bit 16; 0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc_major(int bf16, int n, int a_mn, int b_mn) {
  return make_idesc(bf16, n) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// image column of the concatenated input [x1 | pad | x2 | pad] -> row of W, or -1
__host__ __device__ __forceinline__ int wrow_of_col(int c, int k1, int k1p, int k2) {
  if (c < k1) return c;
  if (c >= k1p && c - k1p < k2) return k1 + (c - k1p);
  return -1;
}

// ---------------------------------------------------------------- weight images
// fwd:   img (nn = output column < Np, kk = input image column < Kp)  = W[wrow(kk)][nn]
// dgrad: img (nn = input image column < Kp, kk = output column < Nz)  = W[wrow(nn)][kk]
// layout [kk/8][NN][8] (swizzle-free K-major operand B), NN = number of nn values.
template <int BF16>
__global__ void wimg_kernel(const float* __restrict__ w, int n, int k1, int k1p, int k2, int Kp,
                            int Nz, int dgrad, uint16_t* __restrict__ img) {
  const int NN = dgrad ? Kp : Nz;
  const int KK = dgrad ? Nz : Kp;
  const int total = NN * KK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i / NN, nn = i % NN;
    float v = 0.f;
    if (dgrad) {
      const int r = wrow_of_col(nn, k1, k1p, k2);
      if (r >= 0 && kk < n) v = w[(size_t)r * n + kk];
    } else {
      const int r = wrow_of_col(kk, k1, k1p, k2);
      if (r >= 0 && nn < n) v = w[(size_t)r * n + nn];
    }
    const uint32_t pk = pack2<BF16, 0>(v, 0.f);
    img[((size_t)(kk / 8) * NN + nn) * 8 + (kk % 8)] = (uint16_t)(pk & 0xFFFFu);
  }
}

// ---------------------------------------------------------------- dz = dy * act'(y), colsum
// dz16[m][Nz] (Nz = n rounded up to 16, pad columns zero); part[block][n] = column sums of the
// fp32 products over the block's rows.
// IN16 = 1 (chain mode): y and dy are 16-bit rows y16[m][ldy] / dy16[m][lddy] (dz16 may alias dy16).
template <int BF16, int IN16>
__global__ void __launch_bounds__(256) act_bwd_colsum_kernel(
    const float* __restrict__ y, const float* __restrict__ dy, long long m, int n, int Nz, int act,
    long long rows_per_block, uint16_t* dz16, float* __restrict__ part,
    const uint16_t* __restrict__ y16, int ldy, const uint16_t* dy16, int lddy) {
  __shared__ float4 red[256];
  const int n4 = n >> 2;
  const int rp = 256 / n4;                       // rows per pass
  const int cq = threadIdx.x % n4, rl = threadIdx.x / n4;
  const bool active = rl < rp;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(m, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    for (long long r = r0 + rl; r < r1; r += rp) {
      float4 yy, dd;
      if (IN16) {
        const uint2 a = *reinterpret_cast<const uint2*>(y16 + r * ldy + cq * 4);
        const uint2 b = *reinterpret_cast<const uint2*>(dy16 + r * lddy + cq * 4);
        yy = make_float4(unpack_lo<BF16>(a.x), unpack_hi<BF16>(a.x), unpack_lo<BF16>(a.y), unpack_hi<BF16>(a.y));
        dd = make_float4(unpack_lo<BF16>(b.x), unpack_hi<BF16>(b.x), unpack_lo<BF16>(b.y), unpack_hi<BF16>(b.y));
      } else {
        yy = *reinterpret_cast<const float4*>(y + r * n + cq * 4);
        dd = *reinterpret_cast<const float4*>(dy + r * n + cq * 4);
      }
      float4 z;
      z.x = dd.x * act_grad_tc(act, yy.x);
      z.y = dd.y * act_grad_tc(act, yy.y);
      z.z = dd.z * act_grad_tc(act, yy.z);
      z.w = dd.w * act_grad_tc(act, yy.w);
      acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
      uint2 o;
      o.x = pack2<BF16, 0>(z.x, z.y);
      o.y = pack2<BF16, 0>(z.z, z.w);
      *reinterpret_cast<uint2*>(dz16 + r * Nz + cq * 4) = o;
      if (cq == n4 - 1)
        for (int c = n; c < Nz; c += 4) *reinterpret_cast<uint2*>(dz16 + r * Nz + c) = make_uint2(0u, 0u);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < n4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < rp; ++j) {
      const float4 v = red[j * n4 + threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(part + (size_t)blockIdx.x * n + threadIdx.x * 4) = s;
  }
}

// out[i] += sum_p part[p][i]   (fixed order: deterministic).  A block owns 32 consecutive
// outputs; its 8 thread slices each sum every 8th partial (128-byte coalesced rows), the slices are
// combined in a fixed order through shared memory.  (Round 1's one-thread-per-output loop over all
// partials took 18 us per layer for 148 x 16 K floats -- 10 % of the configs[3] train step.)
constexpr int RP_OUT = 32, RP_SLICES = 8;
__global__ void __launch_bounds__(RP_OUT * RP_SLICES)
reduce_partials_kernel(const float* __restrict__ part, int nparts, int count, int stride,
                       float* __restrict__ out) {
  __shared__ float red[RP_SLICES][RP_OUT];
  const int lane = threadIdx.x % RP_OUT, slice = threadIdx.x / RP_OUT;
  const int i = blockIdx.x * RP_OUT + lane;
  float s = 0.f;
  if (i < count)
    for (int p = slice; p < nparts; p += RP_SLICES) s += part[(size_t)p * stride + i];
  red[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && i < count) {
    float t = red[0][lane];
#pragma unroll
    for (int j = 1; j < RP_SLICES; ++j) t += red[j][lane];
    out[i] += t;
  }
}

// ---------------------------------------------------------------- row GEMM (fwd / dgrad)
struct RowGemmParams {
  const float* x1; int k1, ld1;        // A16 = 0: fp32 inputs [rows][ld]
  const float* x2; int k2, ld2;
  int k1p;                             // image column where x2 starts (k1 rounded up to 16)
  const uint16_t* a16; int lda16;      // A16 = 1: 16-bit input [rows][lda16], Kp columns used
  const uint16_t* wimg;                // [Kp/8][Np][8]
  int Kp, Np;
  const float* bias; int nbias; int act;
  float* out1; int n1, ldo1;           // image columns [0, n1) -> out1
  float* out2; int n2, ldo2, col2;     // image columns [col2, col2 + n2) -> out2
  uint16_t* out16; int ldo16;          // O16 = 1: image columns [0, n1) -> 16-bit rows instead of out1
  int stages;                          // A-operand images in flight (2..4, as shared memory allows)
  long long rows;
  long long tiles;
};
__device__ __forceinline__ void stg256u(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y),
               "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.256) halve the LSU wavefronts of the one-row-per-lane
// access pattern these kernels use
__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y),
               "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}

// 8 consecutive image columns c .. c+7 of the concatenated fp32 input row -> two float4
// (zeros outside x1 / x2; regions start on 8-column boundaries, lengths are multiples of 4)
struct RowSrc {
  const float* s1; const float* s2;
  int k1, k1p, k2;
  bool v1, v2;                 // rows of x1 / x2 are 32-byte aligned
};
__device__ __forceinline__ void load8(const RowSrc& r, int c, float4& a, float4& b) {
  a = make_float4(0.f, 0.f, 0.f, 0.f);
  b = a;
  if (c < r.k1) {
    if (c + 8 <= r.k1 && r.v1) ldg256(r.s1 + c, a, b);
    else {
      a = __ldg(reinterpret_cast<const float4*>(r.s1 + c));
      if (c + 8 <= r.k1) b = __ldg(reinterpret_cast<const float4*>(r.s1 + c + 4));
    }
  } else if (c >= r.k1p && c - r.k1p < r.k2) {
    const int d = c - r.k1p;
    if (d + 8 <= r.k2 && r.v2) ldg256(r.s2 + d, a, b);
    else {
      a = __ldg(reinterpret_cast<const float4*>(r.s2 + d));
      if (d + 8 <= r.k2) b = __ldg(reinterpret_cast<const float4*>(r.s2 + d + 4));
    }
  }
}
template <int BF16>
__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) {
  uint4 q;
  q.x = pack2<BF16, 0>(a.x, a.y);
  q.y = pack2<BF16, 0>(a.z, a.w);
  q.z = pack2<BF16, 0>(b.x, b.y);
  q.w = pack2<BF16, 0>(b.z, b.w);
  return q;
}

// warp 0 MMA, warps 1-8 loaders (two threads per tile row), warps 9-16 epilogue (two per
// TMEM lane quarter)
constexpr int RG_THREADS = 544;

// O16 = 1 (chain mode): the result columns [0, n1) are stored as 16-bit rows (out16), rounded the
// way the next layer's operand image would round them anyway.
template <int BF16, int A16, int O16>
__global__ void __launch_bounds__(RG_THREADS, 1) rowgemm_tc_kernel(const RowGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int Kp = p.Kp, Np = p.Np;
  const uint32_t wbytes = (uint32_t)Kp * Np * 2;
  const uint32_t abytes = (uint32_t)Kp * 128 * 2;
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + wbytes;
  const int S = p.stages;         // the loaders run up to S - 1 tiles ahead of the tensor pipe
  float* s_bias = reinterpret_cast<float*>(s_a + (size_t)S * abytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* a_full = bars;        // [4] 256 loader arrivals
  uint64_t* a_empty = bars + 4;   // [4] commit
  uint64_t* d_full = bars + 8;    // [2] commit
  uint64_t* d_empty = bars + 10;  // [2] 256 epilogue arrivals
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(a_full + i, 256); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(d_full + i, 1); mbar_init(d_empty + i, 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weight image + bias
  for (uint32_t i = threadIdx.x; i < wbytes / 16; i += RG_THREADS)
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  for (int i = threadIdx.x; i < 256; i += RG_THREADS)
    s_bias[i] = (p.bias && i < p.nbias) ? p.bias[i] : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long ntile = (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == 0) {
    // ============================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BF16, Np);
      const uint32_t a0 = smem_u32(s_a), w0 = smem_u32(s_w);
      const uint32_t lbo_b = (uint32_t)Np * 16;
      for (long long it = 0; it < ntile; ++it) {
        const int buf = (int)(it & 1);
        const int sa = (int)(it % S);
        mbar_wait(a_full + sa, (uint32_t)(it / S) & 1);
        if (it >= 2) mbar_wait(d_empty + buf, (uint32_t)((it >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t d_t = tmem_base + buf * 256;
        for (int ks = 0; ks < Kp / 16; ++ks)
          tc_mma_ss(d_t, make_b_desc(a0 + sa * abytes + ks * 2 * 2048, 2048, 128),
                    make_b_desc(w0 + ks * 2 * lbo_b, lbo_b, 128), idesc, ks > 0 ? 1u : 0u);
        tc_commit(a_empty + sa);
        tc_commit(d_full + buf);
      }
    }
  } else if (warp <= 8) {
    // ================================================================= loaders
    const int lt = threadIdx.x - 32;
    const int t = lt & 127, half = lt >> 7;
    const int KG = Kp / 8, KGh = (KG + 1) / 2;
    const int kg_lo = half * KGh, kg_hi = min(KG, kg_lo + KGh);
    RowSrc rs;
    rs.k1 = p.k1; rs.k1p = p.k1p; rs.k2 = p.k2;
    rs.v1 = !A16 && (p.ld1 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.x1) & 31) == 0);
    rs.v2 = !A16 && p.x2 && (p.ld2 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.x2) & 31) == 0);
    // 16-bit rows: a warp instruction covers 8 rows x 4 sixteen-byte chunks (64 contiguous bytes per
    // row from global memory, 128 contiguous bytes per quarter warp into the image).  With <= 16
    // chunks per row the loads of tile it + 1 are issued BEFORE the image of tile it is written: the
    // global-memory latency of one tile hides behind the previous one (it was paid once per tile).
    const int lw = lt >> 5, r8 = lane & 7, kq = lane >> 3;
    auto load16 = [&](long long it, int kg0, uint4 (&v)[2][4]) {
      const long long tile_row0 = (it * gridDim.x + blockIdx.x) * 128;
#pragma unroll
      for (int rsub = 0; rsub < 2; ++rsub) {
        const long long grow = tile_row0 + lw * 16 + rsub * 8 + r8;
        const uint16_t* src = p.a16 + grow * p.lda16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kg = kg0 + 4 * j + kq;
          v[rsub][j] = make_uint4(0u, 0u, 0u, 0u);
          if (grow < p.rows && kg < KG) v[rsub][j] = __ldg(reinterpret_cast<const uint4*>(src + kg * 8));
        }
      }
    };
    auto store16 = [&](uint8_t* imgb, int kg0, const uint4 (&v)[2][4]) {
#pragma unroll
      for (int rsub = 0; rsub < 2; ++rsub) {
        const int tr = lw * 16 + rsub * 8 + r8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kg = kg0 + 4 * j + kq;
          if (kg < KG) *reinterpret_cast<uint4*>(imgb + (size_t)kg * 2048 + (size_t)tr * 16) = v[rsub][j];
        }
      }
    };
    if (A16 && KG <= 16) {
      uint4 va[2][4], vb[2][4];
      if (ntile > 0) load16(0, 0, va);
      for (long long it = 0; it < ntile; it += 2) {
        // even tile from va (prefetch odd into vb), odd tile from vb (prefetch next even into va)
        if (it + 1 < ntile) load16(it + 1, 0, vb);
        {
          const int buf = (int)(it % S);
          if (it >= S) mbar_wait(a_empty + buf, (uint32_t)((it / S) - 1) & 1);
          store16(s_a + (size_t)buf * abytes, 0, va);
          fence_proxy_async();
          mbar_arrive(a_full + buf);
        }
        if (it + 1 < ntile) {
          if (it + 2 < ntile) load16(it + 2, 0, va);
          const long long i1 = it + 1;
          const int buf = (int)(i1 % S);
          if (i1 >= S) mbar_wait(a_empty + buf, (uint32_t)((i1 / S) - 1) & 1);
          store16(s_a + (size_t)buf * abytes, 0, vb);
          fence_proxy_async();
          mbar_arrive(a_full + buf);
        }
      }
    } else
    for (long long it = 0; it < ntile; ++it) {
      const int buf = (int)(it % S);
      const long long row = (it * gridDim.x + blockIdx.x) * 128 + t;
      const bool valid = row < p.rows;
      if (it >= S) mbar_wait(a_empty + buf, (uint32_t)((it / S) - 1) & 1);
      uint8_t* img = s_a + (size_t)buf * abytes + (size_t)t * 16;
      if (A16) {
        uint8_t* imgb = s_a + (size_t)buf * abytes;
        for (int kg0 = 0; kg0 < KG; kg0 += 16) {
          uint4 v[2][4];
          load16(it, kg0, v);
          store16(imgb, kg0, v);
        }
      } else {
        rs.s1 = p.x1 + row * p.ld1;
        rs.s2 = p.x2 ? p.x2 + row * p.ld2 : nullptr;
        for (int kg0 = kg_lo; kg0 < kg_hi; kg0 += 8) {
          float4 va[8], vb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[j] = va[j];
            if (valid && kg0 + j < kg_hi) load8(rs, (kg0 + j) * 8, va[j], vb[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (kg0 + j < kg_hi)
              *reinterpret_cast<uint4*>(img + (size_t)(kg0 + j) * 2048) = pack8<BF16>(va[j], vb[j]);
        }
      }
      fence_proxy_async();
      mbar_arrive(a_full + buf);
    }
  } else {
    // ================================================================ epilogue
    const int wq = warp & 3;
    const int ch = (warp - 9) >> 2;
    const int t = wq * 32 + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool w1 = p.out1 && (p.ldo1 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out1) & 31) == 0);
    const bool w2 = p.out2 && (p.ldo2 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out2) & 31) == 0);
    // The activation is a run-time argument but fixed for the launch: the tile loop is instantiated
    // once per kind (with `apply_act(p.act, ..)` per element the compiler inlined the sigmoid and
    // softplus code 64 times per tile row behind uniform branches -- 1 300 SASS instructions of
    // epilogue, which bounded the kernel at ~9 k clocks per 128-row tile; profiles/r2_rowgemm.md).
    auto run = [&](auto act_c) {
      constexpr int ACT = decltype(act_c)::value;
      for (long long it = 0; it < ntile; ++it) {
        const int buf = (int)(it & 1);
        const long long row = (it * gridDim.x + blockIdx.x) * 128 + t;
        const bool valid = row < p.rows;
        mbar_wait(d_full + buf, (uint32_t)(it >> 1) & 1);
        tc_fence_after();
        for (int c0 = ch * 16; c0 < Np; c0 += 32) {
          uint32_t r[16];
          TC_LD16(r, tb + buf * 256 + c0);
          float4 bb[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) bb[q] = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * q);
          tc_wait_ld();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int c = c0 + 8 * j;
              float4 o[2];
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const float4 bq = bb[2 * j + q];
                o[q].x = apply_act(ACT, __uint_as_float(r[8 * j + 4 * q + 0]) + bq.x);
                o[q].y = apply_act(ACT, __uint_as_float(r[8 * j + 4 * q + 1]) + bq.y);
                o[q].z = apply_act(ACT, __uint_as_float(r[8 * j + 4 * q + 2]) + bq.z);
                o[q].w = apply_act(ACT, __uint_as_float(r[8 * j + 4 * q + 3]) + bq.w);
              }
              if (O16 && c + 8 <= p.n1) {
                *reinterpret_cast<uint4*>(p.out16 + row * p.ldo16 + c) = pack8<BF16>(o[0], o[1]);
              } else if (c + 8 <= p.n1) {
                if (w1) stg256(p.out1 + row * p.ldo1 + c, o[0], o[1]);
                else if (p.out1) {
                  *reinterpret_cast<float4*>(p.out1 + row * p.ldo1 + c) = o[0];
                  *reinterpret_cast<float4*>(p.out1 + row * p.ldo1 + c + 4) = o[1];
                }
              } else if (c + 4 <= p.n1) {
                if (p.out1) *reinterpret_cast<float4*>(p.out1 + row * p.ldo1 + c) = o[0];
              } else if (p.out2 && c >= p.col2 && c < p.col2 + p.n2) {
                const int d = c - p.col2;
                if (d + 8 <= p.n2 && w2) stg256(p.out2 + row * p.ldo2 + d, o[0], o[1]);
                else {
                  *reinterpret_cast<float4*>(p.out2 + row * p.ldo2 + d) = o[0];
                  if (d + 8 <= p.n2) *reinterpret_cast<float4*>(p.out2 + row * p.ldo2 + d + 4) = o[1];
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(d_empty + buf);
      }
    };
    switch (p.act) {
      case NF_ACT_RELU: run(std::integral_constant<int, NF_ACT_RELU>{}); break;
      case NF_ACT_SIGMOID: run(std::integral_constant<int, NF_ACT_SIGMOID>{}); break;
      case NF_ACT_SOFTPLUS: run(std::integral_constant<int, NF_ACT_SOFTPLUS>{}); break;
      default: run(std::integral_constant<int, NF_ACT_NONE>{}); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// ---------------------------------------------------------------- weight gradient
struct WgradParams {
  const float* x1; int k1, ld1;
  const float* x2; int k2, ld2;
  int k1p, Kp;                         // image columns of [x1 | pad | x2 | pad]; Kp multiple of 16
  int KpW;                             // Kp rounded up to 128 (MMA M blocks)
  const uint16_t* dz16; int Nz;        // [rows][Nz]
  int n;                               // real output columns
  float* part;                         // [gridDim.x][(k1 + k2) * n]
  long long rows, tiles;
  const uint16_t* x16; int ldx16;      // X16 = 1: the input as 16-bit rows covering the Kp image columns
};

constexpr int WG_THREADS = 288;        // warp 0 MMA, warps 1-8 loaders (+ final epilogue)

template <int BF16, int X16>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int KpW = p.KpW, Nz = p.Nz;
  const uint32_t xbytes = (uint32_t)KpW * 128 * 2, zbytes = (uint32_t)Nz * 128 * 2;
  uint8_t* s_x = smem;                              // [2][KpW/8][128 rows][8]
  uint8_t* s_z = smem + 2 * xbytes;                 // [2][Nz/8][128 rows][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_z + 2 * zbytes);
  uint64_t* full = bars;          // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* done = bars + 4;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(full + i, 256); mbar_init(empty + i, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero both X images once: pad feature groups stay zero for the whole kernel
  for (uint32_t i = threadIdx.x; i < 2 * xbytes / 16; i += WG_THREADS)
    reinterpret_cast<uint4*>(s_x)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const long long ntile = (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int mblocks = KpW / 128;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_major(BF16, Nz, 1, 1);
      // MN-major, swizzle-free (cute UMMA canonical layout ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO))):
      // 8 (MN) x 8 (K) core matrix with MN contiguous (16 B) and K rows 16 B apart.  Here K =
      // tile row: row groups are 128 B apart (LBO), feature groups 2048 B apart (SBO).
      const uint32_t lbo = 128u, sbo = 2048u;
      const uint32_t x0 = smem_u32(s_x), z0 = smem_u32(s_z);
      for (long long it = 0; it < ntile; ++it) {
        const int buf = (int)(it & 1);
        mbar_wait(full + buf, (uint32_t)(it >> 1) & 1);
        tc_fence_after();
        for (int mb = 0; mb < mblocks; ++mb)
          for (int ks = 0; ks < 8; ++ks)
            tc_mma_ss(tmem_base + mb * 256,
                      make_b_desc(x0 + buf * xbytes + mb * 16 * 2048 + ks * 256, lbo, sbo),
                      make_b_desc(z0 + buf * zbytes + ks * 256, lbo, sbo), idesc,
                      (it > 0 || ks > 0) ? 1u : 0u);
        tc_commit(empty + buf);
      }
      tc_commit(done);
    }
  } else {
    const int lt = threadIdx.x - 32;
    const int t = lt & 127, half = lt >> 7;
    const int KG = p.Kp / 8, KGh = (KG + 1) / 2;
    const int kg_lo = half * KGh, kg_hi = min(KG, kg_lo + KGh);
    const int ZG = Nz / 8, ZGh = (ZG + 1) / 2;
    const int zg_lo = half * ZGh, zg_hi = min(ZG, zg_lo + ZGh);
    RowSrc rs;
    rs.k1 = p.k1; rs.k1p = p.k1p; rs.k2 = p.k2;
    rs.v1 = (p.ld1 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.x1) & 31) == 0);
    rs.v2 = p.x2 && (p.ld2 % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.x2) & 31) == 0);
    // 16-bit operands (X16): plain copies, 8 rows x 4 sixteen-byte chunks per warp instruction (64
    // contiguous bytes per row from global memory); with <= 16 chunks per row in both operands the
    // loads of tile it + 1 are in flight while tile it is written to shared memory.
    const int lw = lt >> 5, r8 = lane & 7, kq = lane >> 3;
    auto wload = [&](long long it, const uint16_t* base, int ld, int G, int g0, uint4 (&v)[2][4]) {
      const long long tile_row0 = (it * gridDim.x + blockIdx.x) * 128;
#pragma unroll
      for (int rsub = 0; rsub < 2; ++rsub) {
        const long long grow = tile_row0 + lw * 16 + rsub * 8 + r8;
        const uint16_t* src = base + grow * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kg = g0 + 4 * j + kq;
          v[rsub][j] = make_uint4(0u, 0u, 0u, 0u);
          if (grow < p.rows && kg < G) v[rsub][j] = __ldg(reinterpret_cast<const uint4*>(src + kg * 8));
        }
      }
    };
    auto wstore = [&](uint8_t* img, int G, int g0, const uint4 (&v)[2][4]) {
#pragma unroll
      for (int rsub = 0; rsub < 2; ++rsub) {
        const int tr = lw * 16 + rsub * 8 + r8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kg = g0 + 4 * j + kq;
          if (kg < G) *reinterpret_cast<uint4*>(img + (size_t)kg * 2048 + (size_t)tr * 16) = v[rsub][j];
        }
      }
    };
    if (X16 && KG <= 16 && ZG <= 16) {
      uint4 xa[2][4], za[2][4], xb[2][4], zb2[2][4];
      if (ntile > 0) { wload(0, p.x16, p.ldx16, KG, 0, xa); wload(0, p.dz16, Nz, ZG, 0, za); }
      for (long long it = 0; it < ntile; it += 2) {
        if (it + 1 < ntile) { wload(it + 1, p.x16, p.ldx16, KG, 0, xb); wload(it + 1, p.dz16, Nz, ZG, 0, zb2); }
        {
          const int buf = (int)(it & 1);
          if (it >= 2) mbar_wait(empty + buf, (uint32_t)((it >> 1) - 1) & 1);
          wstore(s_x + (size_t)buf * xbytes, KG, 0, xa);
          wstore(s_z + (size_t)buf * zbytes, ZG, 0, za);
          fence_proxy_async();
          mbar_arrive(full + buf);
        }
        if (it + 1 < ntile) {
          if (it + 2 < ntile) { wload(it + 2, p.x16, p.ldx16, KG, 0, xa); wload(it + 2, p.dz16, Nz, ZG, 0, za); }
          const long long i1 = it + 1;
          const int buf = (int)(i1 & 1);
          if (i1 >= 2) mbar_wait(empty + buf, (uint32_t)((i1 >> 1) - 1) & 1);
          wstore(s_x + (size_t)buf * xbytes, KG, 0, xb);
          wstore(s_z + (size_t)buf * zbytes, ZG, 0, zb2);
          fence_proxy_async();
          mbar_arrive(full + buf);
        }
      }
    } else
    for (long long it = 0; it < ntile; ++it) {
      const int buf = (int)(it & 1);
      const long long row = (it * gridDim.x + blockIdx.x) * 128 + t;
      const bool valid = row < p.rows;
      if (it >= 2) mbar_wait(empty + buf, (uint32_t)((it >> 1) - 1) & 1);
      if (X16) {
        uint8_t* xb = s_x + (size_t)buf * xbytes;
        uint8_t* zb = s_z + (size_t)buf * zbytes;
        for (int g0 = 0; g0 < KG; g0 += 16) {
          uint4 v[2][4];
          wload(it, p.x16, p.ldx16, KG, g0, v);
          wstore(xb, KG, g0, v);
        }
        for (int g0 = 0; g0 < ZG; g0 += 16) {
          uint4 v[2][4];
          wload(it, p.dz16, Nz, ZG, g0, v);
          wstore(zb, ZG, g0, v);
        }
        fence_proxy_async();
        mbar_arrive(full + buf);
        continue;
      }
      uint8_t* xi = s_x + (size_t)buf * xbytes + (size_t)t * 16;
      uint8_t* zi = s_z + (size_t)buf * zbytes + (size_t)t * 16;
      rs.s1 = p.x1 + row * p.ld1;
      rs.s2 = p.x2 ? p.x2 + row * p.ld2 : nullptr;
      const uint16_t* zs = p.dz16 + row * Nz;
      // dz chunks first (short), then the X groups
      uint4 zv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        zv[j] = make_uint4(0u, 0u, 0u, 0u);
        if (valid && zg_lo + j < zg_hi) zv[j] = __ldg(reinterpret_cast<const uint4*>(zs + (zg_lo + j) * 8));
      }
      for (int kg0 = kg_lo; kg0 < kg_hi; kg0 += 8) {
        float4 va[8], vb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          vb[j] = va[j];
          if (valid && kg0 + j < kg_hi) load8(rs, (kg0 + j) * 8, va[j], vb[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (kg0 + j < kg_hi)
            *reinterpret_cast<uint4*>(xi + (size_t)(kg0 + j) * 2048) = pack8<BF16>(va[j], vb[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (zg_lo + j < zg_hi) *reinterpret_cast<uint4*>(zi + (size_t)(zg_lo + j) * 2048) = zv[j];
      for (int g0 = zg_lo + 8; g0 < zg_hi; g0 += 8) {             // Nz = 256 only
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (valid && g0 + j < zg_hi) v = __ldg(reinterpret_cast<const uint4*>(zs + (g0 + j) * 8));
          if (g0 + j < zg_hi) *reinterpret_cast<uint4*>(zi + (size_t)(g0 + j) * 2048) = v;
        }
      }
      fence_proxy_async();
      mbar_arrive(full + buf);
    }
    // ---- final: accumulators -> this CTA's partial (warps 1-4: one per TMEM lane quarter)
    if (warp <= 4) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int wq = warp & 3;
      const int tl = wq * 32 + lane;                     // TMEM lane = image column within block
      const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
      float* part = p.part + (size_t)blockIdx.x * (size_t)(p.k1 + p.k2) * p.n;
      for (int mb = 0; mb < mblocks; ++mb) {
        const int wr = wrow_of_col(mb * 128 + tl, p.k1, p.k1p, p.k2);
        for (int c0 = 0; c0 < Nz; c0 += 16) {
          uint32_t r[16];
          TC_LD16(r, tb + mb * 256 + c0);
          tc_wait_ld();
          if (wr >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c0 + 4 * j;
              if (c + 4 <= p.n)
                *reinterpret_cast<float4*>(part + (size_t)wr * p.n + c) =
                    make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

inline int rup(int x, int m) { return (x + m - 1) / m * m; }
inline size_t rup256(size_t x) { return (x + 255) / 256 * 256; }

struct TcDims {
  int k1p, k2p, Kp, KpW, Nz;
};
inline TcDims tc_dims(int k1, int k2, int n) {
  TcDims d;
  d.k1p = k2 ? rup(k1, 16) : rup(k1, 16);
  d.k2p = k2 ? rup(k2, 16) : 0;
  d.Kp = d.k1p + d.k2p;
  d.KpW = rup(d.Kp, 128);
  d.Nz = rup(n, 16);
  return d;
}
constexpr size_t RG_SMEM_EXTRA = 256 * 4 + 16 * 8;
// shared memory of rowgemm_tc_kernel and how many A-operand images fit (2..4): weight image
// wk x wn, A images of ak columns (x 128 rows x 2 B)
inline size_t rg_smem(int wk, int wn, int ak, int* stages) {
  const size_t wbytes = (size_t)wk * wn * 2, ab = (size_t)ak * 256;
  long long S = ((long long)200 * 1024 - (long long)wbytes - (long long)RG_SMEM_EXTRA) / (long long)ab;
  S = S < 2 ? 2 : (S > 4 ? 4 : S);
  *stages = (int)S;
  return wbytes + (size_t)S * ab + RG_SMEM_EXTRA;
}
constexpr int COLSUM_BLOCKS = 592;

}  // namespace

// Shapes the tensor-core path handles; everything else stays on the FP32 CUDA-core kernels.
bool nf_dense_tc_supported(int k1, int k2, int n) {
  const TcDims d = tc_dims(k1, k2, n);
  if (n % 4 || k1 % 4 || k2 % 4 || n < 4) return false;
  if (d.Kp > 256 || d.Nz > 256) return false;
  const size_t fwd = (size_t)d.Kp * d.Nz * 2 + 2 * (size_t)d.Kp * 256 + RG_SMEM_EXTRA;
  const size_t dgr = (size_t)d.Kp * d.Nz * 2 + 2 * (size_t)d.Nz * 256 + RG_SMEM_EXTRA;
  const size_t wgr = 2 * (size_t)(d.KpW + d.Nz) * 256 + 128;
  const size_t lim = 220 * 1024;
  return fwd <= lim && dgr <= lim && wgr <= lim;
}

size_t nf_dense_tc_fwd_workspace(int k1, int k2, int n) {
  const TcDims d = tc_dims(k1, k2, n);
  return rup256((size_t)d.Kp * d.Nz * 2);
}

size_t nf_dense_tc_bwd_workspace(long long m, int k1, int k2, int n, int sm_count) {
  const TcDims d = tc_dims(k1, k2, n);
  size_t b = rup256((size_t)m * d.Nz * 2);                       // dz16
  b += rup256((size_t)d.Kp * d.Nz * 2);                          // dgrad weight image
  b += rup256((size_t)COLSUM_BLOCKS * n * 4);                    // colsum partials
  b += rup256((size_t)sm_count * (size_t)(k1 + k2) * n * 4);     // wgrad partials
  return b;
}

template <int BF16>
static int dense_tc_fwd_t(nf_ctx* ctx, const float* x1, int k1, const float* x2, int k2,
                          const float* w, const float* b, long long m, int n, int act, float* y,
                          void* work, cudaStream_t st) {
  const TcDims d = tc_dims(k1, k2, n);
  uint16_t* img = reinterpret_cast<uint16_t*>(work);
  wimg_kernel<BF16><<<64, 256, 0, st>>>(w, n, k1, d.k1p, k2, d.Kp, d.Nz, 0, img);
  NF_LAUNCH_CHECK(ctx);
  RowGemmParams p;
  memset(&p, 0, sizeof(p));
  p.x1 = x1; p.k1 = k1; p.ld1 = k1; p.x2 = x2; p.k2 = k2; p.ld2 = k2; p.k1p = d.k1p;
  p.wimg = img; p.Kp = d.Kp; p.Np = d.Nz; p.bias = b; p.nbias = n; p.act = act;
  p.out1 = y; p.n1 = n; p.ldo1 = n;
  p.rows = m; p.tiles = (m + 127) / 128;
  const size_t smb = rg_smem(d.Kp, d.Nz, d.Kp, &p.stages);
  const int grid = (int)std::min<long long>(ctx->sm_count, p.tiles);
  NF_CUDA(ctx, cudaFuncSetAttribute(rowgemm_tc_kernel<BF16, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
  rowgemm_tc_kernel<BF16, 0, 0><<<grid, RG_THREADS, smb, st>>>(p);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

template <int BF16>
static int dense_tc_bwd_t(nf_ctx* ctx, const float* x1, int k1, const float* x2, int k2,
                          const float* w, const float* y, const float* dy, long long m, int n,
                          int act, float* dx1, float* dx2, float* dw, float* db, void* work,
                          cudaStream_t st) {
  const TcDims d = tc_dims(k1, k2, n);
  uint8_t* wp = reinterpret_cast<uint8_t*>(work);
  uint16_t* dz16 = reinterpret_cast<uint16_t*>(wp);
  wp += rup256((size_t)m * d.Nz * 2);
  uint16_t* img = reinterpret_cast<uint16_t*>(wp);
  wp += rup256((size_t)d.Kp * d.Nz * 2);
  float* cpart = reinterpret_cast<float*>(wp);
  wp += rup256((size_t)COLSUM_BLOCKS * n * 4);
  float* wpart = reinterpret_cast<float*>(wp);
  // dz (16 bit) + bias gradient
  const long long rpb = (m + COLSUM_BLOCKS - 1) / COLSUM_BLOCKS;
  const int cblocks = (int)((m + rpb - 1) / rpb);
  act_bwd_colsum_kernel<BF16, 0><<<cblocks, 256, 0, st>>>(y, dy, m, n, d.Nz, act, rpb, dz16, cpart, nullptr, 0, nullptr, 0);
  NF_LAUNCH_CHECK(ctx);
  if (db) {
    reduce_partials_kernel<<<(n + RP_OUT - 1) / RP_OUT, RP_OUT * RP_SLICES, 0, st>>>(cpart, cblocks, n, n, db);
    NF_LAUNCH_CHECK(ctx);
  }
  const long long tiles = (m + 127) / 128;
  const int grid = (int)std::min<long long>(ctx->sm_count, tiles);
  if (dx1 || dx2) {
    wimg_kernel<BF16><<<64, 256, 0, st>>>(w, n, k1, d.k1p, k2, d.Kp, d.Nz, 1, img);
    NF_LAUNCH_CHECK(ctx);
    RowGemmParams p;
    memset(&p, 0, sizeof(p));
    p.a16 = dz16; p.lda16 = d.Nz; p.wimg = img;
    p.Kp = d.Nz;                       // contraction over the output features
    p.Np = d.Kp;                       // result columns = input image columns
    p.act = NF_ACT_NONE;
    p.out1 = dx1; p.n1 = k1; p.ldo1 = k1;
    p.out2 = dx2; p.n2 = k2; p.ldo2 = k2; p.col2 = d.k1p;
    p.rows = m; p.tiles = tiles;
    const size_t smb = rg_smem(d.Kp, d.Nz, d.Nz, &p.stages);
    NF_CUDA(ctx, cudaFuncSetAttribute(rowgemm_tc_kernel<BF16, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    rowgemm_tc_kernel<BF16, 1, 0><<<grid, RG_THREADS, smb, st>>>(p);
    NF_LAUNCH_CHECK(ctx);
  }
  if (dw) {
    WgradParams q;
    memset(&q, 0, sizeof(q));
    q.x1 = x1; q.k1 = k1; q.ld1 = k1; q.x2 = x2; q.k2 = k2; q.ld2 = k2;
    q.k1p = d.k1p; q.Kp = d.Kp; q.KpW = d.KpW; q.dz16 = dz16; q.Nz = d.Nz; q.n = n;
    q.part = wpart; q.rows = m; q.tiles = tiles;
    const size_t smb = 2 * (size_t)(d.KpW + d.Nz) * 256 + 128;
    NF_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel<BF16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    wgrad_tc_kernel<BF16, 0><<<grid, WG_THREADS, smb, st>>>(q);
    NF_LAUNCH_CHECK(ctx);
    const int count = (k1 + k2) * n;
    reduce_partials_kernel<<<(count + RP_OUT - 1) / RP_OUT, RP_OUT * RP_SLICES, 0, st>>>(wpart, grid, count, count, dw);
    NF_LAUNCH_CHECK(ctx);
  }
  return NF_OK;
}

int nf_dense_tc_fwd(nf_ctx* ctx, const float* x1, int k1, const float* x2, int k2, const float* w,
                    const float* b, long long m, int n, int act, float* y, void* work,
                    int precision, cudaStream_t st) {
  return precision == NF_PREC_BF16
             ? dense_tc_fwd_t<1>(ctx, x1, k1, x2, k2, w, b, m, n, act, y, work, st)
             : dense_tc_fwd_t<0>(ctx, x1, k1, x2, k2, w, b, m, n, act, y, work, st);
}

int nf_dense_tc_bwd(nf_ctx* ctx, const float* x1, int k1, const float* x2, int k2, const float* w,
                    const float* y, const float* dy, long long m, int n, int act, float* dx1,
                    float* dx2, float* dw, float* db, void* work, int precision, cudaStream_t st) {
  return precision == NF_PREC_BF16
             ? dense_tc_bwd_t<1>(ctx, x1, k1, x2, k2, w, y, dy, m, n, act, dx1, dx2, dw, db, work, st)
             : dense_tc_bwd_t<0>(ctx, x1, k1, x2, k2, w, y, dy, m, n, act, dx1, dx2, dw, db, work, st);
}


// =====================================================================================
// Whole-network forward / backward for the train step (SURVEY 8b "*_bwd counterparts"): one C call
// runs every Dense of an mlp.Network (nerfactor/networks/mlp.py:39-50 + the seq.Network head) and
// keeps the activations between the layers -- and between forward and backward -- as 16-BIT rows
// in a caller-provided workspace.  The operand images of the per-layer kernels round activations
// to 16 bit anyway, so the numbers match the layer-by-layer path (nf_dense_fwd / nf_dense_bwd with
// fp32 activations) up to the rounding of the bias-gradient column sums; the HBM traffic per
// hidden layer drops from 4 + 4 to 2 + 2 bytes per activation.
namespace {

template <int BF16>
__global__ void cvt16_kernel(const float* __restrict__ x, long long rows, int k, int kp,
                             uint16_t* __restrict__ d1, int ld1, uint16_t* __restrict__ d2, int ld2,
                             int off2) {
  const int groups = kp / 8;
  const long long total = rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c = (int)(i % groups) * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c + j < k) ? x[r * k + c + j] : 0.f;
    uint4 q;
    q.x = pack2<BF16, 0>(v[0], v[1]); q.y = pack2<BF16, 0>(v[2], v[3]);
    q.z = pack2<BF16, 0>(v[4], v[5]); q.w = pack2<BF16, 0>(v[6], v[7]);
    *reinterpret_cast<uint4*>(d1 + r * ld1 + c) = q;
    if (d2) *reinterpret_cast<uint4*>(d2 + r * ld2 + off2 + c) = q;
  }
}

__global__ void axpy_kernel(const float* __restrict__ a, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] += a[i];
}

struct ChainPlan {
  int depth, in_dim, k0p, skip;               // skip = index of the layer fed with [h | x], 0 = none
  int width[NF_CHAIN_MAX], nz[NF_CHAIN_MAX];  // output columns, rounded up to 16
  int ld[NF_CHAIN_MAX];                       // row stride of the 16-bit output buffer of hidden layer l
  size_t off_x0, off_h[NF_CHAIN_MAX];         // saved activations
  size_t off_dh[2], off_dzh, off_img, off_cpart, off_wpart, off_dxs, total;
};

bool chain_plan(const nf_mlp_chain* c, long long rows, int sm_count, ChainPlan& P) {
  if (!c || c->depth < 2 || c->depth > NF_CHAIN_MAX || c->in_dim < 4 || c->in_dim % 4) return false;
  P.depth = c->depth; P.in_dim = c->in_dim; P.k0p = rup(c->in_dim, 16); P.skip = c->skip_layer;
  if (P.skip < 0 || P.skip >= c->depth) return false;
  size_t off = 0;
  P.off_x0 = off; off += rup256((size_t)rows * P.k0p * 2);
  int max_k = P.k0p, max_n = 16;
  for (int l = 0; l < c->depth; ++l) {
    P.width[l] = c->width[l];
    P.nz[l] = rup(c->width[l], 16);
    const bool hidden = l < c->depth - 1;
    if (hidden && (c->width[l] % 16 || c->width[l] > 256)) return false;      // 16-bit rows need whole chunks
    if (!hidden && c->width[l] % 4) return false;
    P.ld[l] = P.nz[l] + ((P.skip && l == P.skip - 1) ? P.k0p : 0);
    P.off_h[l] = off;
    if (hidden) off += rup256((size_t)rows * P.ld[l] * 2);
    const int kin = l == 0 ? P.k0p : P.ld[l - 1];
    if (kin > 256 || P.nz[l] > 256) return false;
    max_k = std::max(max_k, kin); max_n = std::max(max_n, P.nz[l]);
  }
  P.off_dh[0] = off; off += rup256((size_t)rows * 256 * 2);
  P.off_dh[1] = off; off += rup256((size_t)rows * 256 * 2);
  P.off_dzh = off; off += rup256((size_t)rows * P.nz[c->depth - 1] * 2);
  P.off_img = off; off += rup256((size_t)max_k * max_n * 2);
  P.off_cpart = off; off += rup256((size_t)COLSUM_BLOCKS * 256 * 4);
  P.off_wpart = off; off += rup256((size_t)sm_count * (size_t)max_k * max_n * 4);
  P.off_dxs = off; off += rup256((size_t)rows * c->in_dim * 4);
  P.total = off;
  return true;
}

template <int BF16>
int chain_fwd_t(nf_ctx* ctx, const nf_mlp_chain* c, const ChainPlan& P, const float* x,
                long long rows, float* y, uint8_t* ws, cudaStream_t st) {
  uint16_t* x0 = reinterpret_cast<uint16_t*>(ws + P.off_x0);
  uint16_t* hskip = P.skip ? reinterpret_cast<uint16_t*>(ws + P.off_h[P.skip - 1]) : nullptr;
  const long long tiles = (rows + 127) / 128;
  const int grid = (int)std::min<long long>(ctx->sm_count, tiles);
  cvt16_kernel<BF16><<<(unsigned)std::min<long long>((rows * (P.k0p / 8) + 255) / 256, 4096), 256, 0, st>>>(
      x, rows, P.in_dim, P.k0p, x0, P.k0p, hskip, P.skip ? P.ld[P.skip - 1] : 0,
      P.skip ? P.nz[P.skip - 1] : 0);
  NF_LAUNCH_CHECK(ctx);
  uint16_t* img = reinterpret_cast<uint16_t*>(ws + P.off_img);
  for (int l = 0; l < P.depth; ++l) {
    const bool hidden = l < P.depth - 1;
    const bool sk = P.skip && l == P.skip;
    const int k1 = l == 0 ? P.in_dim : P.width[l - 1];
    const int k2 = sk ? P.in_dim : 0;
    const TcDims d = tc_dims(k1, k2, P.width[l]);
    wimg_kernel<BF16><<<64, 256, 0, st>>>(c->w[l], P.width[l], k1, d.k1p, k2, d.Kp, d.Nz, 0, img);
    NF_LAUNCH_CHECK(ctx);
    RowGemmParams p;
    memset(&p, 0, sizeof(p));
    p.a16 = l == 0 ? x0 : reinterpret_cast<uint16_t*>(ws + P.off_h[l - 1]);
    p.lda16 = l == 0 ? P.k0p : P.ld[l - 1];
    p.k1p = d.k1p; p.wimg = img; p.Kp = d.Kp; p.Np = d.Nz;
    p.bias = c->b[l]; p.nbias = P.width[l]; p.act = c->act[l];
    p.n1 = P.width[l];
    p.rows = rows; p.tiles = tiles;
    const size_t smb = rg_smem(d.Kp, d.Nz, d.Kp, &p.stages);
    if (hidden) {
      p.out16 = reinterpret_cast<uint16_t*>(ws + P.off_h[l]); p.ldo16 = P.ld[l];
      NF_CUDA(ctx, cudaFuncSetAttribute(rowgemm_tc_kernel<BF16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
      rowgemm_tc_kernel<BF16, 1, 1><<<grid, RG_THREADS, smb, st>>>(p);
    } else {
      p.out1 = y; p.ldo1 = P.width[l];
      NF_CUDA(ctx, cudaFuncSetAttribute(rowgemm_tc_kernel<BF16, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
      rowgemm_tc_kernel<BF16, 1, 0><<<grid, RG_THREADS, smb, st>>>(p);
    }
    NF_LAUNCH_CHECK(ctx);
  }
  return NF_OK;
}

template <int BF16>
int chain_bwd_t(nf_ctx* ctx, const nf_mlp_chain* c, const ChainPlan& P, long long rows,
                const float* y, const float* dy, float* dx, float* const* dw, float* const* db,
                uint8_t* ws, cudaStream_t st) {
  const long long tiles = (rows + 127) / 128;
  const int grid = (int)std::min<long long>(ctx->sm_count, tiles);
  uint16_t* x0 = reinterpret_cast<uint16_t*>(ws + P.off_x0);
  uint16_t* img = reinterpret_cast<uint16_t*>(ws + P.off_img);
  float* cpart = reinterpret_cast<float*>(ws + P.off_cpart);
  float* wpart = reinterpret_cast<float*>(ws + P.off_wpart);
  float* dxs = reinterpret_cast<float*>(ws + P.off_dxs);
  const long long rpb = (rows + COLSUM_BLOCKS - 1) / COLSUM_BLOCKS;
  const int cblocks = (int)((rows + rpb - 1) / rpb);
  int cur = 0;                                      // dh[cur]: gradient w.r.t. the output of layer l
  for (int l = P.depth - 1; l >= 0; --l) {
    const bool hidden = l < P.depth - 1;
    const bool sk = P.skip && l == P.skip;
    const int n = P.width[l];
    const int k1 = l == 0 ? P.in_dim : P.width[l - 1];
    const int k2 = sk ? P.in_dim : 0;
    const TcDims d = tc_dims(k1, k2, n);
    // dz = dy * act'(y) as 16 bit (+ bias gradient): hidden layers in place in dh[cur]
    uint16_t* dz16;
    if (hidden) {
      dz16 = reinterpret_cast<uint16_t*>(ws + P.off_dh[cur]);
      const uint16_t* h = reinterpret_cast<const uint16_t*>(ws + P.off_h[l]);
      act_bwd_colsum_kernel<BF16, 1><<<cblocks, 256, 0, st>>>(nullptr, nullptr, rows, n, d.Nz, c->act[l], rpb,
                                                              dz16, cpart, h, P.ld[l], dz16, d.Nz);
    } else {
      dz16 = reinterpret_cast<uint16_t*>(ws + P.off_dzh);
      act_bwd_colsum_kernel<BF16, 0><<<cblocks, 256, 0, st>>>(y, dy, rows, n, d.Nz, c->act[l], rpb, dz16,
                                                              cpart, nullptr, 0, nullptr, 0);
    }
    NF_LAUNCH_CHECK(ctx);
    if (db && db[l]) {
      reduce_partials_kernel<<<(n + RP_OUT - 1) / RP_OUT, RP_OUT * RP_SLICES, 0, st>>>(cpart, cblocks, n, n, db[l]);
      NF_LAUNCH_CHECK(ctx);
    }
    // weight gradient: [x1 | x2]^T dz with the 16-bit input rows of the forward pass
    if (dw && dw[l]) {
      WgradParams q;
      memset(&q, 0, sizeof(q));
      q.k1 = k1; q.k2 = k2; q.k1p = d.k1p; q.Kp = d.Kp; q.KpW = d.KpW; q.dz16 = dz16; q.Nz = d.Nz; q.n = n;
      q.x16 = l == 0 ? x0 : reinterpret_cast<const uint16_t*>(ws + P.off_h[l - 1]);
      q.ldx16 = l == 0 ? P.k0p : P.ld[l - 1];
      q.part = wpart; q.rows = rows; q.tiles = tiles;
      const size_t smb = 2 * (size_t)(d.KpW + d.Nz) * 256 + 128;
      NF_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel<BF16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
      wgrad_tc_kernel<BF16, 1><<<grid, WG_THREADS, smb, st>>>(q);
      NF_LAUNCH_CHECK(ctx);
      const int count = (k1 + k2) * n;
      reduce_partials_kernel<<<(count + RP_OUT - 1) / RP_OUT, RP_OUT * RP_SLICES, 0, st>>>(wpart, grid, count, count, dw[l]);
      NF_LAUNCH_CHECK(ctx);
    }
    // data gradient: dz W^T -> 16-bit rows for the layer below (fp32 for the network input)
    if (l > 0 || dx) {
      wimg_kernel<BF16><<<64, 256, 0, st>>>(c->w[l], n, k1, d.k1p, k2, d.Kp, d.Nz, 1, img);
      NF_LAUNCH_CHECK(ctx);
      RowGemmParams p;
      memset(&p, 0, sizeof(p));
      p.a16 = dz16; p.lda16 = d.Nz; p.wimg = img;
      p.Kp = d.Nz; p.Np = d.Kp; p.act = NF_ACT_NONE;
      p.rows = rows; p.tiles = tiles;
      const size_t smb = rg_smem(d.Kp, d.Nz, d.Nz, &p.stages);
      if (l == 0) {
        p.out1 = dx; p.n1 = k1; p.ldo1 = k1;
        NF_CUDA(ctx, cudaFuncSetAttribute(rowgemm_tc_kernel<BF16, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
        rowgemm_tc_kernel<BF16, 1, 0><<<grid, RG_THREADS, smb, st>>>(p);
      } else {
        p.out16 = reinterpret_cast<uint16_t*>(ws + P.off_dh[cur ^ 1]); p.ldo16 = P.nz[l - 1]; p.n1 = k1;
        if (sk && dx) { p.out2 = dxs; p.n2 = k2; p.ldo2 = k2; p.col2 = d.k1p; }
        NF_CUDA(ctx, cudaFuncSetAttribute(rowgemm_tc_kernel<BF16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
        rowgemm_tc_kernel<BF16, 1, 1><<<grid, RG_THREADS, smb, st>>>(p);
      }
      NF_LAUNCH_CHECK(ctx);
      cur ^= 1;
    }
  }
  if (dx && P.skip) {                                  // dx += the skip layer's input part
    const long long total = rows * P.in_dim;
    axpy_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 8192), 256, 0, st>>>(dxs, dx, total);
    NF_LAUNCH_CHECK(ctx);
  }
  return NF_OK;
}

}  // namespace

extern "C" {

size_t nf_mlp_chain_workspace_bytes(const nf_mlp_chain* chain, long long rows) {
  ChainPlan P;
  if (rows <= 0 || !chain_plan(chain, rows, 160, P)) return 0;
  return P.total;
}

int nf_mlp_chain_fwd(nf_ctx* ctx, const nf_mlp_chain* chain, const float* x_d, long long rows,
                     float* y_d, void* workspace_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, chain && rows >= 0, "bad argument");
  NF_CHECK_ARG(ctx, precision == NF_PREC_F16 || precision == NF_PREC_BF16,
               "nf_mlp_chain_fwd: tensor-core precisions only (NF_PREC_F16 / NF_PREC_BF16)");
  if (rows == 0) return NF_OK;
  ChainPlan P;
  if (!chain_plan(chain, rows, 160, P))
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "nf_mlp_chain: hidden widths must be multiples of 16 and <= 256, layer inputs "
                        "<= 256 columns, 2 <= depth <= 8");
  NF_CHECK_ARG(ctx, x_d && y_d && workspace_d, "null buffer");
  for (int l = 0; l < chain->depth; ++l) NF_CHECK_ARG(ctx, chain->w[l] && chain->b[l], "null weights");
  NF_CHECK_ARG(ctx, ctx->sm_count <= 160, "unexpected SM count");
  return precision == NF_PREC_BF16
             ? chain_fwd_t<1>(ctx, chain, P, x_d, rows, y_d, (uint8_t*)workspace_d, (cudaStream_t)stream)
             : chain_fwd_t<0>(ctx, chain, P, x_d, rows, y_d, (uint8_t*)workspace_d, (cudaStream_t)stream);
}

int nf_mlp_chain_bwd(nf_ctx* ctx, const nf_mlp_chain* chain, long long rows, const float* y_d,
                     const float* dy_d, float* dx_d, float* const* dw_d, float* const* db_d,
                     void* workspace_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, chain && rows >= 0, "bad argument");
  NF_CHECK_ARG(ctx, precision == NF_PREC_F16 || precision == NF_PREC_BF16,
               "nf_mlp_chain_bwd: tensor-core precisions only (NF_PREC_F16 / NF_PREC_BF16)");
  if (rows == 0) return NF_OK;
  ChainPlan P;
  if (!chain_plan(chain, rows, 160, P))
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED, "nf_mlp_chain: unsupported network shape");
  NF_CHECK_ARG(ctx, y_d && dy_d && workspace_d, "null buffer");
  NF_CHECK_ARG(ctx, ctx->sm_count <= 160, "unexpected SM count");
  return precision == NF_PREC_BF16
             ? chain_bwd_t<1>(ctx, chain, P, rows, y_d, dy_d, dx_d, dw_d, db_d, (uint8_t*)workspace_d,
                              (cudaStream_t)stream)
             : chain_bwd_t<0>(ctx, chain, P, rows, y_d, dy_d, dx_d, dw_d, db_d, (uint8_t*)workspace_d,
                              (cudaStream_t)stream);
}

}  // extern "C"
