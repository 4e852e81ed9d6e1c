// tcgen05 forward of the per-point networks (normal / albedo / BRDF-latent or roughness:
// nerfactor/models/shape.py:196-211, nerfactor.py:377-411) at ~fp32 accuracy (NF_PREC_F16X3).
//
// These nets feed RGB multiplicatively, so plain fp16 operands (2^-11) are not accurate enough
// (DESIGN.md section 5).  Every operand is split into an fp16 "hi" part and an fp16 "lo"
// residual, x = hi + lo with |lo| <= 2^-11 |x|, and each Dense layer is three MMA chains
//     D = A_hi W_hi + A_lo W_hi + A_hi W_lo          (the lo*lo term is < 2^-22 and dropped)
// accumulated in fp32 in TMEM: ~2^-21 relative accuracy at 3x the (tiny) tensor work.
// Structure = the sigma kernel's, simplified: persistent CTA, one 128-point tile in flight,
// hi/lo activations in TMEM, hi/lo positional encoding as shared-memory A operands, the
// 256 KB of hi/lo weights streamed through a 5 x 32 KB ring with cp.async.bulk in consumption
// order, fp32 bias / ReLU / head on the CUDA cores.
#include "nf_common.cuh"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nftc;

constexpr int PT_THREADS = 256;               // warp 0 MMA, warp 1 producer, warps 4-7 workers
constexpr int PT_NSLOT = 5;
constexpr int PT_SLOT_BYTES = 32768;
constexpr int PT_E_BYTES = 16384;             // 128 rows x 64 k x 2 B
constexpr int PCOL_D = 0, PCOL_AHI = 128, PCOL_ALO = 192;
constexpr uint32_t PT_LBO = 128 * 16, PT_SBO = 128;
constexpr int PT_NCHUNK = 10;                 // see chunk table below

constexpr size_t PT_OFF_RING = 0;
constexpr size_t PT_OFF_E = PT_OFF_RING + (size_t)PT_NSLOT * PT_SLOT_BYTES;   // E_hi, E_lo
constexpr size_t PT_OFF_AUX = PT_OFF_E + 2 * PT_E_BYTES;     // bias[4][128], wout[128][4], bout[4]
constexpr size_t PT_AUX_FLOATS = 4 * 128 + 128 * 4 + 4;
constexpr size_t PT_OFF_BAR = PT_OFF_AUX + PT_AUX_FLOATS * 4;
constexpr size_t PT_SMEM = PT_OFF_BAR + 24 * 8;

// Weight chunks in consumption order.  kind 0: hidden K-block (128 k, 32 KB), kind 1: input
// block (64 k, 16 KB); lo = 0: W_hi (multiplied by A_hi and A_lo), lo = 1: W_lo (A_hi only).
//   L0: (1,0) (1,1)   L1: (0,0) (0,1)   L2: (0,0) (0,1)   L3: (0,0) (0,1) (1,0) (1,1)
__device__ __forceinline__ void chunk_info(int c, int& layer, int& kind, int& lo) {
  const int tl[PT_NCHUNK] = {0, 0, 1, 1, 2, 2, 3, 3, 3, 3};
  const int tk[PT_NCHUNK] = {1, 1, 0, 0, 0, 0, 0, 0, 1, 1};
  layer = tl[c]; kind = tk[c]; lo = c & 1;
}

struct PointTcParams {
  const uint8_t* blob;
  size_t off_img, off_aux;
  const float* xyz;      // [n,3]
  float xyz_scale;
  int n, out_dim, out_act;
  int tiles_per_cta;
  float* out;            // [n, out_dim]
};

// split x into fp16 hi / lo pairs: returns packed hi (two values) and writes packed lo
__device__ __forceinline__ uint32_t split2(float x0, float x1, uint32_t& lo) {
  __half2 h = __floats2half2_rn(x0, x1);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  lo = *reinterpret_cast<uint32_t*>(&l);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(PT_THREADS, 1) point_tc_kernel(const PointTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_ring = smem + PT_OFF_RING;
  uint8_t* s_e = smem + PT_OFF_E;
  const float* s_bias = reinterpret_cast<const float*>(smem + PT_OFF_AUX);
  const float* s_wout = s_bias + 4 * 128;       // [128][4]
  const float* s_bout = s_wout + 128 * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PT_OFF_BAR);
  uint64_t* bar_full = bars;            // [5]
  uint64_t* bar_empty = bars + 5;       // [5]
  uint64_t* bar_d = bars + 10;          // accumulator ready
  uint64_t* bar_a = bars + 11;          // A operand (TMEM hi/lo or smem E) ready, 128 arrivals
  uint64_t* bar_w = bars + 12;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = p.tiles_per_cta;

  if (threadIdx.x == 0) {
    for (int i = 0; i < PT_NSLOT; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, 1); }
    mbar_init(bar_d, 1); mbar_init(bar_a, 128); mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_w, (uint32_t)(PT_AUX_FLOATS * 4));
    bulk_g2s(smem + PT_OFF_AUX, p.blob + p.off_aux, (uint32_t)(PT_AUX_FLOATS * 4), bar_w);
  }
  mbar_wait(bar_w, 0);

  if (warp == 1) {
    // ============================================================ TMA producer
    if (lane == 0) {
      const uint8_t* img = p.blob + p.off_img;
      uint32_t fill = 0;
      for (int it = 0; it < ntile; ++it) {
        uint32_t off = 0;
        for (int c = 0; c < PT_NCHUNK; ++c) {
          int layer, kind, lo;
          chunk_info(c, layer, kind, lo);
          const uint32_t bytes = kind ? 16384u : 32768u;
          const uint32_t slot = fill % PT_NSLOT;
          if (fill >= PT_NSLOT) mbar_wait(bar_empty + slot, ((fill / PT_NSLOT) - 1) & 1);
          mbar_expect_tx(bar_full + slot, bytes);
          bulk_g2s(s_ring + (size_t)slot * PT_SLOT_BYTES, img + off, bytes, bar_full + slot);
          off += bytes;
          ++fill;
        }
      }
    }
  } else if (warp == 0) {
    // ============================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(0, 128);
      const uint32_t ring0 = smem_u32(s_ring), e0 = smem_u32(s_e);
      uint32_t fill = 0, na = 0;
      for (int it = 0; it < ntile; ++it) {
        int cur_layer = -1;
        for (int c = 0; c < PT_NCHUNK; ++c) {
          int layer, kind, lo;
          chunk_info(c, layer, kind, lo);
          if (layer != cur_layer) {            // this layer's A operand is complete
            mbar_wait(bar_a, na & 1);
            ++na;
            tc_fence_after();
            cur_layer = layer;
          }
          const uint32_t slot = fill % PT_NSLOT;
          mbar_wait(bar_full + slot, (fill / PT_NSLOT) & 1);
          tc_fence_after();
          const uint32_t b0 = ring0 + slot * PT_SLOT_BYTES;
          const uint32_t d_t = tmem_base + PCOL_D;
          const bool first = (c == 0) || (layer > 0 && kind == 0 && lo == 0);
          if (kind == 1) {
            // input block from shared memory: E_hi (and E_lo for the W_hi chunk)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_ss(d_t, make_b_desc(e0 + ks * 2 * PT_LBO, PT_LBO, PT_SBO),
                        make_b_desc(b0 + ks * 2 * PT_LBO, PT_LBO, PT_SBO), idesc,
                        (first && ks == 0) ? 0u : 1u);
            if (!lo) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc_mma_ss(d_t, make_b_desc(e0 + PT_E_BYTES + ks * 2 * PT_LBO, PT_LBO, PT_SBO),
                          make_b_desc(b0 + ks * 2 * PT_LBO, PT_LBO, PT_SBO), idesc, 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              tc_mma_ts(d_t, tmem_base + PCOL_AHI + ks * 8,
                        make_b_desc(b0 + ks * 2 * PT_LBO, PT_LBO, PT_SBO), idesc,
                        (first && ks == 0) ? 0u : 1u);
            if (!lo) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                tc_mma_ts(d_t, tmem_base + PCOL_ALO + ks * 8,
                          make_b_desc(b0 + ks * 2 * PT_LBO, PT_LBO, PT_SBO), idesc, 1u);
            }
          }
          tc_commit(bar_empty + slot);
          ++fill;
          // last chunk of a layer: L0 -> c = 1, L1 -> 3, L2 -> 5, L3 -> 9
          if (c == 1 || c == 3 || c == 5 || c == 9) tc_commit(bar_d);
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================ workers
    const int wq = warp - 4;
    const int t = wq * 32 + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t nd = 0;
    for (int it = 0; it < ntile; ++it) {
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      // ---- positional encoding of this point -> E_hi / E_lo (embedder.py:46-47)
      {
        float v[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = 0.f;
        if (g < p.n) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float pc = p.xyz[g * 3 + c] * p.xyz_scale;
            v[c] = pc;
#pragma unroll
            for (int f = 0; f < 10; ++f) {
              float s, co;
              sincosf(pc * (float)(1 << f), &s, &co);
              v[3 + 6 * f + c] = s;
              v[3 + 6 * f + 3 + c] = co;
            }
          }
        }
        // E is free: the previous tile's last reader (layer-3 MMAs) completed before bar_d
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 qh, ql;
          qh.x = split2(v[8 * j + 0], v[8 * j + 1], ql.x);
          qh.y = split2(v[8 * j + 2], v[8 * j + 3], ql.y);
          qh.z = split2(v[8 * j + 4], v[8 * j + 5], ql.z);
          qh.w = split2(v[8 * j + 6], v[8 * j + 7], ql.w);
          *reinterpret_cast<uint4*>(s_e + ((size_t)j * 128 + t) * 16) = qh;
          *reinterpret_cast<uint4*>(s_e + PT_E_BYTES + ((size_t)j * 128 + t) * 16) = ql;
        }
        fence_proxy_async();
        mbar_arrive(bar_a);
      }
      // ---- layers 0..2: D -> bias + ReLU -> (hi, lo) -> TMEM
      for (int l = 0; l < 3; ++l) {
        mbar_wait(bar_d, nd & 1);
        ++nd;
        tc_fence_after();
        const float* bias = s_bias + l * 128;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t r0[32], r1[32];
          TC_LD32(r0, tb + PCOL_D + c2 * 64);
          TC_LD32(r1, tb + PCOL_D + c2 * 64 + 32);
          tc_wait_ld();
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 bb = *reinterpret_cast<const float2*>(bias + c2 * 64 + 2 * i);
            ph[i] = split2(fmaxf(__uint_as_float(r0[2 * i]) + bb.x, 0.f),
                           fmaxf(__uint_as_float(r0[2 * i + 1]) + bb.y, 0.f), pl[i]);
          }
          TC_ST16(tb + PCOL_AHI + c2 * 32, ph);
          TC_ST16(tb + PCOL_ALO + c2 * 32, pl);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 bb = *reinterpret_cast<const float2*>(bias + c2 * 64 + 32 + 2 * i);
            ph[i] = split2(fmaxf(__uint_as_float(r1[2 * i]) + bb.x, 0.f),
                           fmaxf(__uint_as_float(r1[2 * i + 1]) + bb.y, 0.f), pl[i]);
          }
          TC_ST16(tb + PCOL_AHI + c2 * 32 + 16, ph);
          TC_ST16(tb + PCOL_ALO + c2 * 32 + 16, pl);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_a);
      }
      // ---- layer 3 + head (fp32)
      mbar_wait(bar_d, nd & 1);
      ++nd;
      tc_fence_after();
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        uint32_t r[32];
        TC_LD32(r, tb + PCOL_D + c2 * 32);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float h = fmaxf(__uint_as_float(r[i]) + s_bias[3 * 128 + c2 * 32 + i], 0.f);
          const float4 w = *reinterpret_cast<const float4*>(s_wout + (c2 * 32 + i) * 4);
          o[0] = fmaf(h, w.x, o[0]); o[1] = fmaf(h, w.y, o[1]);
          o[2] = fmaf(h, w.z, o[2]); o[3] = fmaf(h, w.w, o[3]);
        }
      }
      if (g < p.n) {
        for (int j = 0; j < p.out_dim; ++j)
          p.out[g * p.out_dim + j] = apply_act(p.out_act, o[j] + s_bout[j]);
      }
      tc_fence_before();     // D fully read before the next tile's layer-0 MMAs (bar_a arrive)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256)
                 : "memory");
  }
}

uint16_t hb(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
float hf(float f) { return __half2float(__float2half_rn(f)); }

}  // namespace

// Host packing for a per-point network (width 128, depth 4, skip 2, in_dim 63): the ten weight
// chunks in consumption order, each [K/8][128 n][8] fp16 (hi or lo part of W), then the fp32 aux
// block bias[4][128], wout[128][4] (zero padded), bout[4].
int nf_point_tc_pack(nf_mlp* m) {
  const nf_mlp_desc& d = m->d;
  if (d.kind != NF_MLP_POINT || d.width != 128 || d.depth != 4 || d.skip_at != 2 ||
      d.in_dim != 63 || d.n_freqs_a != 10)
    return NF_OK;
  const size_t halves = (size_t)2 * (64 + 128 + 128 + 128 + 64) * 128;
  size_t base = (m->blob.size() + 255) / 256 * 256;
  m->off_tc_f16 = base;
  m->tc_bytes = halves * 2;
  m->off_tc_aux = base + (halves * 2 + 255) / 256 * 256;
  m->tc_aux_bytes = PT_AUX_FLOATS * 4;
  m->blob.resize(m->off_tc_aux + (m->tc_aux_bytes + 255) / 256 * 256, 0);
  uint16_t* img = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tc_f16);
  float* aux = reinterpret_cast<float*>(m->blob.data() + m->off_tc_aux);
  const int tl[PT_NCHUNK] = {0, 0, 1, 1, 2, 2, 3, 3, 3, 3};
  const int tk[PT_NCHUNK] = {1, 1, 0, 0, 0, 0, 0, 0, 1, 1};
  size_t pos = 0;
  for (int c = 0; c < PT_NCHUNK; ++c) {
    const int l = tl[c], kind = tk[c], lo = c & 1;
    const int kk = kind ? 64 : 128;
    const int kreal = kind ? 63 : 128;
    const int r0 = (kind && l == 3) ? 128 : 0;      // skip layer: input rows follow the hidden rows
    for (int k = 0; k < kk; ++k)
      for (int n = 0; n < 128; ++n) {
        float w = k < kreal ? d.W[l][(size_t)(r0 + k) * 128 + n] : 0.f;
        float v = lo ? (w - hf(w)) : w;
        img[pos + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8)] = hb(v);
      }
    pos += (size_t)kk * 128;
  }
  for (int l = 0; l < 4; ++l) memcpy(aux + l * 128, d.b[l], 128 * sizeof(float));
  for (int c = 0; c < 128; ++c)
    for (int j = 0; j < 4; ++j) aux[4 * 128 + c * 4 + j] = j < d.out_dim ? d.W[4][(size_t)c * d.out_dim + j] : 0.f;
  for (int j = 0; j < 4; ++j) aux[4 * 128 + 128 * 4 + j] = j < d.out_dim ? d.b[4][j] : 0.f;
  return NF_OK;
}

int nf_tc_point_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, int n, float xyz_scale,
                       float* out, cudaStream_t st) {
  NF_CHECK_ARG(ctx, m->dev, "network not uploaded (call nf_mlp_upload first)");
  if (m->tc_bytes == 0)
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "no tcgen05 kernel for this point network (need 4 x 128, skip 2, F = 10); "
                        "use NF_PREC_FP32");
  PointTcParams p;
  memset(&p, 0, sizeof(p));
  p.blob = (const uint8_t*)m->dev;
  p.off_img = m->off_tc_f16;
  p.off_aux = m->off_tc_aux;
  p.xyz = xyz; p.xyz_scale = xyz_scale; p.n = n; p.out = out;
  p.out_dim = m->d.out_dim; p.out_act = m->d.out_act;
  const long long tiles = ((long long)n + 127) / 128;
  int grid = ctx->sm_count;
  if (tiles < grid) grid = (int)tiles;
  p.tiles_per_cta = (int)((tiles + grid - 1) / grid);
  NF_CUDA(ctx, cudaFuncSetAttribute(point_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)PT_SMEM));
  point_tc_kernel<<<grid, PT_THREADS, PT_SMEM, st>>>(p);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
