// tcgen05 fused forward of the full NeRF network with viewing directions
// (nerfactor/models/nerf.py:53-71 layout, evaluated as in _eval_nerf_at, nerf.py:254-290):
//   feat  = enc(embed(p))                       8 x 256 ReLU, input re-concatenated after layer 4
//   sigma = sigma_out(feat)                     Dense(1), raw (ReLU is applied when compositing)
//   rgb   = rgb_out(concat(bottleneck(feat), embed(view)))     256 -> 256 (linear) ;
//                                               283 -> 128 ReLU -> 3, raw (sigmoid when compositing)
// Same machinery as sigma_tc_kernel (nf_sigma_tc.cu, CL = 1): P/Q activation buffers and D0/D1
// accumulator halves in TMEM, weights streamed through a 5 x 32 KB ring in consumption order.
// Two more GEMM layers ride on the trunk: layer 8 = bottleneck (no activation), layer 9 = the
// colour hidden layer (N = 128, one accumulator half) whose third K-block is the view-direction
// encoding (SS-mode operand written by the prologue warps next to the position encoding);
// the 128 -> 3 head and the sigma head are contracted on the CUDA cores in the epilogues.
#include "nf_common.cuh"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nftc;

constexpr int NR_THREADS = 448;   // MMA, producer, 8 epilogue warps, 4 prologue warps
constexpr int NR_NSLOT = 5;
constexpr int NR_SLOT_BYTES = 32768;
constexpr int NR_E_BYTES = 16384;              // 128 rows x 64 k x 2 B
constexpr int NR_V_BYTES = 8192;               // 128 rows x 32 k x 2 B
constexpr int NR_LAYERS = 10;                  // 8 trunk + bottleneck + colour hidden
constexpr int NRC_P = 0, NRC_Q = 128, NRC_D0 = 256, NRC_D1 = 384;
constexpr uint32_t NR_LBO = 128 * 16, NR_SBO = 128;

constexpr size_t NR_OFF_RING = 0;
constexpr size_t NR_OFF_E = NR_OFF_RING + (size_t)NR_NSLOT * NR_SLOT_BYTES;
constexpr size_t NR_OFF_V = NR_OFF_E + 2 * NR_E_BYTES;
constexpr size_t NR_OFF_AUX = NR_OFF_V + 2 * NR_V_BYTES;
// aux (fp32): trunk block of nf_sigma_tc_pack  bias[8][256] | w_out[256] | b_out[4]
//             then the colour block            bias8[256] | bias9[128] | w_rgb[128][4] | b_rgb[4]
constexpr int NR_AUX_TRUNK = 8 * 256 + 256 + 4;
constexpr int NR_AUX_RGB = 256 + 128 + 128 * 4 + 4;
constexpr size_t NR_OFF_PART = NR_OFF_AUX + (size_t)(NR_AUX_TRUNK + NR_AUX_RGB) * 4;   // [4][128] f32
constexpr size_t NR_OFF_BAR = NR_OFF_PART + 4 * 128 * 4;
constexpr size_t NR_SMEM = NR_OFF_BAR + 32 * 8;
static_assert(NR_SMEM <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");

struct NerfTcParams {
  const uint8_t* blob;
  size_t off_trunk, off_rgb;     // streamed 16-bit weight images
  size_t off_aux_trunk, off_aux_rgb;
  const float* rayo;
  const float* rayd;
  const float* z;
  long long total;
  int S;
  int tiles_per_cta;
  float* rgbs;                   // [n_rays, S, 4]: raw r, g, b, sigma
};

// number of weight chunks of (layer l, half h) and their sizes
__device__ __forceinline__ int nr_parts(int l) { return l == 0 ? 1 : ((l == 5 || l == 9) ? 3 : 2); }
__device__ __forceinline__ uint32_t nr_part_bytes(int l, int pi) {
  if (l == 0) return 16384u;
  if (pi < 2) return 32768u;
  return l == 5 ? 16384u : 8192u;             // position block (K = 64) / view block (K = 32)
}

template <int BF16>
__global__ void __launch_bounds__(NR_THREADS, 1) nerf_tc_kernel(const NerfTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_ring = smem + NR_OFF_RING;
  uint8_t* s_e = smem + NR_OFF_E;
  uint8_t* s_v = smem + NR_OFF_V;
  const float* s_bias = reinterpret_cast<const float*>(smem + NR_OFF_AUX);
  const float* s_wout = s_bias + 8 * 256;
  const float* s_bout = s_wout + 256;
  const float* s_bias8 = s_bias + NR_AUX_TRUNK;
  const float* s_bias9 = s_bias8 + 256;
  const float* s_wrgb = s_bias9 + 128;        // [128][4]
  const float* s_brgb = s_wrgb + 128 * 4;
  float* s_part = reinterpret_cast<float*>(smem + NR_OFF_PART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NR_OFF_BAR);
  uint64_t* bar_full = bars;            // [5]
  uint64_t* bar_empty = bars + 5;       // [5]
  uint64_t* bar_dfull = bars + 10;      // [2]
  uint64_t* bar_aready = bars + 12;     // [2]
  uint64_t* bar_eready = bars + 14;     // [2]
  uint64_t* bar_efree = bars + 16;      // [2]
  uint64_t* bar_w = bars + 18;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = p.tiles_per_cta;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NR_NSLOT; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_dfull + i, 1); mbar_init(bar_aready + i, 256);
      mbar_init(bar_eready + i, 128); mbar_init(bar_efree + i, 1);
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_w, (NR_AUX_TRUNK + NR_AUX_RGB) * 4);
    bulk_g2s(smem + NR_OFF_AUX, p.blob + p.off_aux_trunk, NR_AUX_TRUNK * 4, bar_w);
    bulk_g2s(smem + NR_OFF_AUX + NR_AUX_TRUNK * 4, p.blob + p.off_aux_rgb, NR_AUX_RGB * 4, bar_w);
  }
  mbar_wait(bar_w, 0);

  if (warp == 1) {
    // ============================================================ TMA producer
    if (lane == 0) {
      uint32_t fill = 0;
      for (int it = 0; it < ntile; ++it) {
        const uint8_t* src = p.blob + p.off_trunk;
        for (int l = 0; l < NR_LAYERS; ++l) {
          if (l == 8) src = p.blob + p.off_rgb;
          const int nh = l == 9 ? 1 : 2;
          for (int h = 0; h < nh; ++h)
            for (int pi = 0; pi < nr_parts(l); ++pi) {
              const uint32_t bytes = nr_part_bytes(l, pi);
              const uint32_t slot = fill % NR_NSLOT;
              if (fill >= NR_NSLOT) mbar_wait(bar_empty + slot, ((fill / NR_NSLOT) - 1) & 1);
              mbar_expect_tx(bar_full + slot, bytes);
              bulk_g2s(s_ring + (size_t)slot * NR_SLOT_BYTES, src, bytes, bar_full + slot);
              src += bytes;
              ++fill;
            }
        }
      }
    }
  } else if (warp == 0) {
    // ============================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BF16, 128);
      const uint32_t ring0 = smem_u32(s_ring), e0 = smem_u32(s_e), v0 = smem_u32(s_v);
      uint32_t fill = 0, na[2] = {0u, 0u};
      auto wait_a = [&](int h) { mbar_wait(bar_aready + h, na[h] & 1); ++na[h]; };
      for (int it = 0; it < ntile; ++it) {
        const int eb = it & 1;
        mbar_wait(bar_eready + eb, (it >> 1) & 1);
        for (int l = 0; l < NR_LAYERS; ++l) {
          const uint32_t xin = tmem_base + ((l & 1) ? NRC_P : NRC_Q);   // layer l >= 1 reads X_l
          const int nh = l == 9 ? 1 : 2;
          for (int h = 0; h < nh; ++h) {
            const uint32_t d_t = tmem_base + (h ? NRC_D1 : NRC_D0);
            for (int pi = 0; pi < nr_parts(l); ++pi) {
              const int part = l == 0 ? 2 : pi;
              // operand / accumulator hazards.  D1 is free at the next tile's layer 0 without a
              // wait: its last reader (layer 8, half 1) was awaited before layer 9 was issued.
              if (l == 0) { if (it > 0 && h == 0) wait_a(0); }
              else if (h == 0 && part < 2) wait_a(part);
              const uint32_t slot = fill % NR_NSLOT;
              mbar_wait(bar_full + slot, (fill / NR_NSLOT) & 1);
              tc_fence_after();
              const uint32_t b0 = ring0 + slot * NR_SLOT_BYTES;
              if (part == 2 && l != 9) {
                const uint32_t a0 = e0 + eb * NR_E_BYTES;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc_mma_ss(d_t, make_b_desc(a0 + ks * 2 * NR_LBO, NR_LBO, NR_SBO),
                            make_b_desc(b0 + ks * 2 * NR_LBO, NR_LBO, NR_SBO), idesc,
                            (l == 0 && ks == 0) ? 0u : 1u);
              } else if (part == 2) {
                const uint32_t a0 = v0 + eb * NR_V_BYTES;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  tc_mma_ss(d_t, make_b_desc(a0 + ks * 2 * NR_LBO, NR_LBO, NR_SBO),
                            make_b_desc(b0 + ks * 2 * NR_LBO, NR_LBO, NR_SBO), idesc, 1u);
              } else {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  tc_mma_ts(d_t, xin + part * 64 + ks * 8,
                            make_b_desc(b0 + ks * 2 * NR_LBO, NR_LBO, NR_SBO), idesc,
                            (part == 0 && ks == 0) ? 0u : 1u);
              }
              tc_commit(bar_empty + slot);
              ++fill;
            }
            tc_commit(bar_dfull + h);
          }
        }
        tc_commit(bar_efree + eb);        // position + view encodings consumed (view: layer 9)
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ================================================================ epilogue
    const int wq = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int t = wq * 32 + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t nd[2] = {0u, 0u};
    for (int it = 0; it < ntile; ++it) {
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      float acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
      for (int l = 0; l < NR_LAYERS; ++l) {
        const uint32_t xout = tb + ((l & 1) ? NRC_Q : NRC_P);   // layer l writes X_{l+1}
        const int nh = l == 9 ? 1 : 2;
        for (int h = 0; h < nh; ++h) {
          mbar_wait(bar_dfull + h, nd[h] & 1);
          ++nd[h];
          tc_fence_after();
          const float* bias = (l < 8 ? s_bias + l * 256 : (l == 8 ? s_bias8 : s_bias9)) + h * 128 + ch * 64;
          uint32_t r0[32], r1[32];
          TC_LD32(r0, tb + (h ? NRC_D1 : NRC_D0) + ch * 64);
          TC_LD32(r1, tb + (h ? NRC_D1 : NRC_D0) + ch * 64 + 32);
          tc_wait_ld();
          if (l == 7) {                       // sigma head on the post-ReLU features
            const float* wo = s_wout + h * 128 + ch * 64;
#pragma unroll
            for (int i = 0; i < 32; ++i)       // same summation order as sigma_tc_kernel
              acc = fmaf(fmaxf(__uint_as_float(r0[i]) + bias[i], 0.f), wo[i], acc);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              acc = fmaf(fmaxf(__uint_as_float(r1[i]) + bias[32 + i], 0.f), wo[32 + i], acc);
          }
          if (l < 9) {
            uint32_t pk[16];
            if (l != 8) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                pk[i] = pack2<BF16, 1>(__uint_as_float(r0[2 * i]) + bias[2 * i],
                                       __uint_as_float(r0[2 * i + 1]) + bias[2 * i + 1]);
            } else {                          // bottleneck: no activation (nerf.py:68)
#pragma unroll
              for (int i = 0; i < 16; ++i)
                pk[i] = pack2<BF16, 0>(__uint_as_float(r0[2 * i]) + bias[2 * i],
                                       __uint_as_float(r0[2 * i + 1]) + bias[2 * i + 1]);
            }
            TC_ST16(xout + h * 64 + ch * 32, pk);
            if (l != 8) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                pk[i] = pack2<BF16, 1>(__uint_as_float(r1[2 * i]) + bias[32 + 2 * i],
                                       __uint_as_float(r1[2 * i + 1]) + bias[32 + 2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                pk[i] = pack2<BF16, 0>(__uint_as_float(r1[2 * i]) + bias[32 + 2 * i],
                                       __uint_as_float(r1[2 * i + 1]) + bias[32 + 2 * i + 1]);
            }
            TC_ST16(xout + h * 64 + ch * 32 + 16, pk);
            tc_wait_st();
          } else {                            // colour head 128 -> 3 on the ReLU'd hidden layer
            const float* wr = s_wrgb + (ch * 64) * 4;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float hv = fmaxf(__uint_as_float(r0[i]) + bias[i], 0.f);
              const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * i);
              cr = fmaf(hv, w4.x, cr); cg = fmaf(hv, w4.y, cg); cb = fmaf(hv, w4.z, cb);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float hv = fmaxf(__uint_as_float(r1[i]) + bias[32 + i], 0.f);
              const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * (32 + i));
              cr = fmaf(hv, w4.x, cr); cg = fmaf(hv, w4.y, cg); cb = fmaf(hv, w4.z, cb);
            }
          }
          tc_fence_before();
          mbar_arrive(bar_aready + h);
        }
      }
      if (ch == 1) {
        s_part[t] = acc; s_part[128 + t] = cr; s_part[256 + t] = cg; s_part[384 + t] = cb;
      }
      named_bar(1, 256);
      if (ch == 0 && g < p.total) {
        float4 o;
        o.x = cr + s_part[128 + t] + s_brgb[0];
        o.y = cg + s_part[256 + t] + s_brgb[1];
        o.z = cb + s_part[384 + t] + s_brgb[2];
        o.w = acc + s_part[t] + s_bout[0];
        *reinterpret_cast<float4*>(p.rgbs + g * 4) = o;       // nerf.py:282 concat([rgb, sigma])
      }
      named_bar(1, 256);
    }
  } else if (warp >= 10) {
    // ================================================================ prologue
    const int t = (warp - 10) * 32 + lane;
    for (int it = 0; it < ntile; ++it) {
      const int eb = it & 1;
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      float v[64], vw[32];
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) vw[i] = 0.f;
      if (g < p.total) {
        const long long ray = g / p.S;
        const float zz = p.z[g];
        float pc[3], vd[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {                        // pts = rayo + rayd * z (nerf.py:162)
          vd[c] = p.rayd[ray * 3 + c];
          pc[c] = __fadd_rn(p.rayo[ray * 3 + c], __fmul_rn(vd[c], zz));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          v[c] = pc[c];
#pragma unroll
          for (int f0 = 0; f0 < 10; f0 += 3) {
            float s, co;
            sincosf(pc[c] * (float)(1 << f0), &s, &co);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (f0 + j < 10) {
                v[3 + 6 * (f0 + j) + c] = s;
                v[3 + 6 * (f0 + j) + 3 + c] = co;
                const float ns = 2.f * s * co, nc = 1.f - 2.f * s * s;
                s = ns; co = nc;
              }
            }
          }
          // view direction, 4 octaves (nerf.py:91-94): views = rayd broadcast (nerf.py:164)
          vw[c] = vd[c];
#pragma unroll
          for (int f0 = 0; f0 < 4; f0 += 3) {
            float s, co;
            sincosf(vd[c] * (float)(1 << f0), &s, &co);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (f0 + j < 4) {
                vw[3 + 6 * (f0 + j) + c] = s;
                vw[3 + 6 * (f0 + j) + 3 + c] = co;
                const float ns = 2.f * s * co, nc = 1.f - 2.f * s * s;
                s = ns; co = nc;
              }
            }
          }
        }
      }
      if (it >= 2) mbar_wait(bar_efree + eb, ((it >> 1) - 1) & 1);
      uint8_t* e = s_e + eb * NR_E_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 q;
        q.x = pack2<BF16, 0>(v[8 * j + 0], v[8 * j + 1]);
        q.y = pack2<BF16, 0>(v[8 * j + 2], v[8 * j + 3]);
        q.z = pack2<BF16, 0>(v[8 * j + 4], v[8 * j + 5]);
        q.w = pack2<BF16, 0>(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(e + ((size_t)j * 128 + t) * 16) = q;
      }
      uint8_t* ev = s_v + eb * NR_V_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q;
        q.x = pack2<BF16, 0>(vw[8 * j + 0], vw[8 * j + 1]);
        q.y = pack2<BF16, 0>(vw[8 * j + 2], vw[8 * j + 3]);
        q.z = pack2<BF16, 0>(vw[8 * j + 4], vw[8 * j + 5]);
        q.w = pack2<BF16, 0>(vw[8 * j + 6], vw[8 * j + 7]);
        *reinterpret_cast<uint4*>(ev + ((size_t)j * 128 + t) * 16) = q;
      }
      fence_proxy_async();
      mbar_arrive(bar_eready + eb);
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

uint16_t nh_bits(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
uint16_t nbf_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}

}  // namespace

// Appends the colour-branch images to a packed NF_MLP_SIGMA network (before nf_mlp_upload):
//   16-bit (fp16 then bf16), chunk order of the kernel: bottleneck (h, kb) 4 x [16 kg][128 n][8];
//   colour hidden kb0, kb1 2 x [16 kg][128 n][8] and the view block [4 kg][128 n][8] (27 of 32 k);
//   fp32 aux: bias8[256] | bias9[128] | w_rgb[128][4] | b_rgb[4].
extern "C" int nf_mlp_attach_rgb(nf_ctx* ctx, nf_mlp* m, const nf_nerf_rgb_desc* r) {
  NF_CHECK_ARG(ctx, m && r, "null argument");
  NF_CHECK_ARG(ctx, m->d.kind == NF_MLP_SIGMA && m->tc_bytes != 0,
               "colour branch needs the 8 x 256 (skip 4, F = 10) sigma network");
  NF_CHECK_ARG(ctx, m->dev == nullptr, "attach before nf_mlp_upload");
  NF_CHECK_ARG(ctx, r->n_freqs_view == 4 && r->hidden == 128, "colour branch must be 283 -> 128 -> 3 (F_view = 4)");
  NF_CHECK_ARG(ctx, r->w_bottleneck && r->b_bottleneck && r->w_rgb0 && r->b_rgb0 && r->w_rgb1 && r->b_rgb1,
               "null weights");
  const size_t halves = (size_t)4 * 128 * 128 + (size_t)2 * 128 * 128 + (size_t)32 * 128;
  size_t base = (m->blob.size() + 255) / 256 * 256;
  m->off_rgb_f16 = base;
  m->off_rgb_bf16 = base + (halves * 2 + 255) / 256 * 256;
  m->off_rgb_aux = m->off_rgb_bf16 + (halves * 2 + 255) / 256 * 256;
  m->blob.resize(m->off_rgb_aux + ((size_t)NR_AUX_RGB * 4 + 255) / 256 * 256, 0);
  uint16_t* i16 = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_rgb_f16);
  uint16_t* ibf = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_rgb_bf16);
  float* aux = reinterpret_cast<float*>(m->blob.data() + m->off_rgb_aux);
  size_t pos = 0;
  auto put = [&](size_t idx, float v) { i16[idx] = nh_bits(v); ibf[idx] = nbf_bits(v); };
  for (int h = 0; h < 2; ++h)
    for (int kb = 0; kb < 2; ++kb) {
      for (int k = 0; k < 128; ++k)
        for (int n = 0; n < 128; ++n)
          put(pos + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8),
              r->w_bottleneck[(size_t)(kb * 128 + k) * 256 + h * 128 + n]);
      pos += (size_t)128 * 128;
    }
  for (int kb = 0; kb < 2; ++kb) {
    for (int k = 0; k < 128; ++k)
      for (int n = 0; n < 128; ++n)
        put(pos + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8), r->w_rgb0[(size_t)(kb * 128 + k) * 128 + n]);
    pos += (size_t)128 * 128;
  }
  for (int k = 0; k < 32; ++k)
    for (int n = 0; n < 128; ++n)
      put(pos + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8),
          k < 27 ? r->w_rgb0[(size_t)(256 + k) * 128 + n] : 0.f);
  memcpy(aux, r->b_bottleneck, 256 * sizeof(float));
  memcpy(aux + 256, r->b_rgb0, 128 * sizeof(float));
  for (int k = 0; k < 128; ++k)
    for (int c = 0; c < 3; ++c) aux[384 + k * 4 + c] = r->w_rgb1[k * 3 + c];
  for (int c = 0; c < 3; ++c) aux[384 + 512 + c] = r->b_rgb1[c];
  return NF_OK;
}

extern "C" int nf_nerf_fwd(nf_ctx* ctx, const nf_mlp* m, const float* rayo_d, const float* rayd_d,
                           const float* z_d, int n_rays, int n_samples, float* rgbs_d,
                           int precision, void* stream) {
  NF_CHECK_ARG(ctx, m, "null network");
  NF_CHECK_ARG(ctx, n_rays >= 0 && n_samples > 0, "bad sizes");
  if (n_rays == 0) return NF_OK;
  NF_CHECK_ARG(ctx, rayo_d && rayd_d && z_d && rgbs_d, "null buffer");
  NF_CHECK_ARG(ctx, m->dev, "network not uploaded (call nf_mlp_upload first)");
  if (m->off_rgb_f16 == 0)
    return nf_set_error(ctx, NF_ERR_INVALID_ARG, "nf_nerf_fwd: no colour branch (nf_mlp_attach_rgb)");
  if (precision != NF_PREC_F16 && precision != NF_PREC_BF16)
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "nf_nerf_fwd: tcgen05 only (NF_PREC_F16 / NF_PREC_BF16); the fp32 path is "
                        "the layer-by-layer nf_dense_fwd chain of the Python mirror");
  NerfTcParams p;
  memset(&p, 0, sizeof(p));
  const bool bf = precision == NF_PREC_BF16;
  p.blob = (const uint8_t*)m->dev;
  p.off_trunk = bf ? m->off_tc_bf16 : m->off_tc_f16;
  p.off_rgb = bf ? m->off_rgb_bf16 : m->off_rgb_f16;
  p.off_aux_trunk = m->off_tc_aux;
  p.off_aux_rgb = m->off_rgb_aux;
  p.rayo = rayo_d; p.rayd = rayd_d; p.z = z_d; p.S = n_samples; p.rgbs = rgbs_d;
  p.total = (long long)n_rays * n_samples;
  const long long tiles = (p.total + 127) / 128;
  int grid = ctx->sm_count;
  if (tiles < grid) grid = (int)tiles;
  p.tiles_per_cta = (int)((tiles + grid - 1) / grid);
  cudaStream_t st = (cudaStream_t)stream;
  if (bf) {
    NF_CUDA(ctx, cudaFuncSetAttribute(nerf_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NR_SMEM));
    nerf_tc_kernel<1><<<grid, NR_THREADS, NR_SMEM, st>>>(p);
  } else {
    NF_CUDA(ctx, cudaFuncSetAttribute(nerf_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NR_SMEM));
    nerf_tc_kernel<0><<<grid, NR_THREADS, NR_SMEM, st>>>(p);
  }
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
