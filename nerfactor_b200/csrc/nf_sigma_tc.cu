// tcgen05 fused forward of the NeRF sigma network (8 x 256 ReLU, input re-concatenated
// after layer 4, Dense(1) + ReLU head): nerfactor/models/nerf.py:53-71 evaluated as in
// eval_sigma_mlp, nerfactor/geometry_from_nerf.py:322-350.  sm_100a only.
//
// Design (DESIGN.md "K5"):
//  * persistent CTA per SM; one 128-sample tile in flight per CTA.  TMEM (512 columns)
//    holds two ping-pong activation buffers P/Q (128 rows x 256 fp16 = 128 columns
//    each) and two accumulator halves D0/D1 (128 x 128 fp32 = 128 columns each).
//  * each layer is issued as two N-halves of two K-blocks; while the tensor pipe
//    works on half 1, the epilogue warps turn half 0 (tcgen05.ld -> bias + ReLU ->
//    fp16 -> tcgen05.st) into the next layer's A operand, and the next layer's first
//    K-block only needs that half -- no bubble between layers.
//  * the 983 KB of weights cannot live in shared memory: they stream from L2 through a
//    5-slot x 32 KB ring with cp.async.bulk (TMA), in exactly the order the MMAs
//    consume them; in a cluster of CL CTAs each CTA fetches 1/CL of every chunk and
//    multicasts it to all of them, dividing the L2 read traffic by CL.
//  * the positional encoding of the next tile is computed by four prologue warps into
//    a double-buffered shared-memory A operand (SS-mode MMA for layer 0 and the skip).
#include "nf_common.cuh"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nftc;

constexpr int SG_THREADS = 448;   // MMA, producer, 8 epilogue warps, 4 prologue warps
constexpr int SG_NSLOT = 5;
constexpr int SG_SLOT_BYTES = 32768;
constexpr int SG_E_BYTES = 16384;              // 128 rows x 64 k x 2 B
constexpr int SG_W = 256, SG_DEPTH = 8, SG_SKIP = 4;
constexpr int COL_P = 0, COL_Q = 128, COL_D0 = 256, COL_D1 = 384;
constexpr uint32_t SG_LBO = 128 * 16, SG_SBO = 128;

// Shared-memory layout of sigma_tc_kernel<.., NSLOT, ESPLIT>.  ESPLIT = 1 ("f16e"): the
// positional encoding is kept as TWO 16-bit images, hi = rn16(e) and lo = rn16(e - hi), and
// layer 0 / the skip layer contract both with the same weight chunk (one extra K = 64 MMA
// block each, +6.7 % tensor work): the 2^-12 rounding of the sin / cos features -- the
// dominant term of the fp16 error on a sharp density field, DESIGN.md section 5 -- goes away.
// The second image pair costs 32 KB, so that mode runs with a 4-slot weight ring.
template <int NSLOT, int ESPLIT>
struct SgLayout {
  static constexpr size_t off_ring = 0;
  static constexpr size_t off_e = off_ring + (size_t)NSLOT * SG_SLOT_BYTES;   // hi images [2]
  static constexpr size_t off_elo = off_e + 2 * SG_E_BYTES;                     // lo images [2]
  static constexpr size_t off_bias = off_elo + (ESPLIT ? 2 * SG_E_BYTES : 0);  // [8][256] f32
  static constexpr size_t off_wout = off_bias + 8 * 256 * 4;                    // [256] f32
  static constexpr size_t off_bout = off_wout + 256 * 4;                        // [4] f32
  static constexpr size_t off_part = off_bout + 16;                             // [128] f32 head partials
  static constexpr size_t off_bar = off_part + 128 * 4;
  // barriers: full[NSLOT] empty[NSLOT] d_full[2] a_ready[2] e_ready[2] e_free[2] bar_w
  static constexpr size_t total = off_bar + 32 * 8;
};
static_assert(SgLayout<5, 0>::total <= 232448 && SgLayout<4, 1>::total <= 232448, "smem budget");

struct SigmaTcParams {
  const uint8_t* blob;
  size_t off_img;        // streamed fp16/bf16 weight image (chunk order, see nf_sigma_tc_pack)
  size_t off_aux;        // fp32: bias[8][256], w_out[256], b_out[4]
  const float* rayo;     // [n_rays,3]
  const float* rayd;     // [n_rays,3]
  const float* z;        // [n_rays,S]
  long long total;       // n_rays * S
  int S;
  int tiles_per_cta;
  float bbox[6];
  int use_bbox;
  float* sigma;          // [n_rays,S]
  int warp_arrive;       // 1: one elected mbarrier arrival per warp instead of one per thread
  long long* dbg;        // NF_SIGMA_DBG: clock64 stamps of CTA 0, tile 2 ([8 layers][2][6]), else NULL
};
__device__ __forceinline__ long long clk64() {
  long long c;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
  return c;
}
#define SG_STAMP(l, h, k)                                                                   \
  do { if (p.dbg && blockIdx.x == 0 && it == 2) p.dbg[((l) * 2 + (h)) * 6 + (k)] = clk64(); } while (0)

// bytes of the weight chunk for (layer, part): part 0/1 = hidden K-blocks, part 2 = input part
__device__ __forceinline__ uint32_t part_bytes(int part) { return part == 2 ? 16384u : 32768u; }

template <int BF16, int CL, int NSLOT, int ESPLIT>
__global__ void __launch_bounds__(SG_THREADS, 1) sigma_tc_kernel(const SigmaTcParams p) {
  using SL = SgLayout<NSLOT, ESPLIT>;
  constexpr int SG_NSLOT = NSLOT;       // shadows the file-level default inside this kernel
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_ring = smem + SL::off_ring;
  uint8_t* s_e = smem + SL::off_e;
  uint8_t* s_elo = smem + SL::off_elo;
  const float* s_bias = reinterpret_cast<const float*>(smem + SL::off_bias);
  const float* s_wout = reinterpret_cast<const float*>(smem + SL::off_wout);
  const float* s_bout = reinterpret_cast<const float*>(smem + SL::off_bout);
  float* s_part = reinterpret_cast<float*>(smem + SL::off_part);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL::off_bar);
  uint64_t* bar_full = bars;            // [<= 5]
  uint64_t* bar_empty = bars + 5;       // [<= 5]
  uint64_t* bar_dfull = bars + 10;      // [2]
  uint64_t* bar_aready = bars + 12;     // [2]
  uint64_t* bar_eready = bars + 14;     // [2]
  uint64_t* bar_efree = bars + 16;      // [2]
  uint64_t* bar_w = bars + 18;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  const uint16_t cta_mask = (uint16_t)((1u << CL) - 1u);
  const int ntile = p.tiles_per_cta;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SG_NSLOT; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, CL); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_dfull + i, 1); mbar_init(bar_aready + i, p.warp_arrive ? 8 : 256);
      mbar_init(bar_eready + i, p.warp_arrive ? 4 : 128); mbar_init(bar_efree + i, 1);
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
  if (CL > 1) cluster_sync_all();      // every CTA's barriers are initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t aux_bytes = 8 * 256 * 4 + 256 * 4 + 16;
    mbar_expect_tx(bar_w, aux_bytes);
    bulk_g2s(smem + SL::off_bias, p.blob + p.off_aux, aux_bytes, bar_w);
  }
  mbar_wait(bar_w, 0);

  if (warp == 1) {
    // ============================================================ TMA producer
    if (lane == 0) {
      const uint8_t* img = p.blob + p.off_img;
      uint32_t fill = 0;
      for (int it = 0; it < ntile; ++it) {
        uint32_t off = 0;
        for (int l = 0; l < SG_DEPTH; ++l)
          for (int h = 0; h < 2; ++h) {
            const int np = l == 0 ? 1 : (l == SG_SKIP + 1 ? 3 : 2);
            for (int pi = 0; pi < np; ++pi) {
              const int part = l == 0 ? 2 : pi;
              const uint32_t bytes = part_bytes(part);
              const uint32_t slot = fill % SG_NSLOT;
              if (fill >= SG_NSLOT) mbar_wait(bar_empty + slot, ((fill / SG_NSLOT) - 1) & 1);
              mbar_expect_tx(bar_full + slot, bytes);
              const uint32_t sl = bytes / CL;
              uint8_t* dst = s_ring + (size_t)slot * SG_SLOT_BYTES + cta_rank * sl;
              const uint8_t* src = img + off + cta_rank * sl;
              if (CL > 1) bulk_g2s_mc(dst, src, sl, bar_full + slot, cta_mask);
              else bulk_g2s(dst, src, sl, bar_full + slot);
              off += bytes;
              ++fill;
            }
          }
      }
    }
  } else if (warp == 0) {
    // ============================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BF16, 128);
      const uint32_t ring0 = smem_u32(s_ring), e0 = smem_u32(s_e), elo0 = smem_u32(s_elo);
      uint32_t fill = 0, na[2] = {0u, 0u};
      auto wait_a = [&](int h) { mbar_wait(bar_aready + h, na[h] & 1); ++na[h]; };
      for (int it = 0; it < ntile; ++it) {
        const int eb = it & 1;
        mbar_wait(bar_eready + eb, (it >> 1) & 1);
        for (int l = 0; l < SG_DEPTH; ++l) {
          const uint32_t xin = tmem_base + ((l & 1) ? COL_P : COL_Q);   // layer l >= 1 reads X_l
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_t = tmem_base + (h ? COL_D1 : COL_D0);
            const int np = l == 0 ? 1 : (l == SG_SKIP + 1 ? 3 : 2);
            for (int pi = 0; pi < np; ++pi) {
              const int part = l == 0 ? 2 : pi;
              // operand / accumulator hazards
              if (l == 0) { if (it > 0) wait_a(h); }
              else if (h == 0 && part < 2) { wait_a(part); SG_STAMP(l, part, 1); }
              const uint32_t slot = fill % SG_NSLOT;
              mbar_wait(bar_full + slot, (fill / SG_NSLOT) & 1);
              tc_fence_after();
              const uint32_t b0 = ring0 + slot * SG_SLOT_BYTES;
              if (part == 2) {
                const uint32_t a0 = e0 + eb * SG_E_BYTES;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc_mma_ss(d_t, make_b_desc(a0 + ks * 2 * SG_LBO, SG_LBO, SG_SBO),
                            make_b_desc(b0 + ks * 2 * SG_LBO, SG_LBO, SG_SBO), idesc,
                            (l == 0 && ks == 0) ? 0u : 1u);
                if (ESPLIT) {          // + lo(e) . W with the same weight chunk
                  const uint32_t al = elo0 + eb * SG_E_BYTES;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    tc_mma_ss(d_t, make_b_desc(al + ks * 2 * SG_LBO, SG_LBO, SG_SBO),
                              make_b_desc(b0 + ks * 2 * SG_LBO, SG_LBO, SG_SBO), idesc, 1u);
                }
              } else {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  tc_mma_ts(d_t, xin + part * 64 + ks * 8,
                            make_b_desc(b0 + ks * 2 * SG_LBO, SG_LBO, SG_SBO), idesc,
                            (part == 0 && ks == 0) ? 0u : 1u);
              }
              if (CL > 1) tc_commit_mc(bar_empty + slot, cta_mask);
              else tc_commit(bar_empty + slot);
              ++fill;
            }
            tc_commit(bar_dfull + h);
            SG_STAMP(l, h, 0);
          }
          if (l == SG_SKIP + 1) tc_commit(bar_efree + eb);   // embedding buffer consumed
        }
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ================================================================ epilogue
    // 8 warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4; every thread
    // converts 64 of the 128 accumulator columns of an N-half per layer.
    const int wq = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int t = wq * 32 + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t nd[2] = {0u, 0u};
    for (int it = 0; it < ntile; ++it) {
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      float acc = 0.f;
      for (int l = 0; l < SG_DEPTH; ++l) {
        const uint32_t xout = tb + ((l & 1) ? COL_Q : COL_P);   // layer l writes X_{l+1}
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_dfull + h, nd[h] & 1);
          ++nd[h];
          tc_fence_after();
          if (threadIdx.x == 64) SG_STAMP(l, h, 2);
          const float* bias = s_bias + l * 256 + h * 128 + ch * 64;
          uint32_t r0[32], r1[32];
          TC_LD32(r0, tb + (h ? COL_D1 : COL_D0) + ch * 64);
          TC_LD32(r1, tb + (h ? COL_D1 : COL_D0) + ch * 64 + 32);
          tc_wait_ld();
          if (threadIdx.x == 64) SG_STAMP(l, h, 3);
          if (l < SG_DEPTH - 1) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 bb = *reinterpret_cast<const float4*>(bias + 4 * i);
              pk[2 * i] = pack2<BF16, 1>(__uint_as_float(r0[4 * i]) + bb.x,
                                             __uint_as_float(r0[4 * i + 1]) + bb.y);
              pk[2 * i + 1] = pack2<BF16, 1>(__uint_as_float(r0[4 * i + 2]) + bb.z,
                                                 __uint_as_float(r0[4 * i + 3]) + bb.w);
            }
            TC_ST16(xout + h * 64 + ch * 32, pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 bb = *reinterpret_cast<const float4*>(bias + 32 + 4 * i);
              pk[2 * i] = pack2<BF16, 1>(__uint_as_float(r1[4 * i]) + bb.x,
                                             __uint_as_float(r1[4 * i + 1]) + bb.y);
              pk[2 * i + 1] = pack2<BF16, 1>(__uint_as_float(r1[4 * i + 2]) + bb.z,
                                                 __uint_as_float(r1[4 * i + 3]) + bb.w);
            }
            TC_ST16(xout + h * 64 + ch * 32 + 16, pk);
            tc_wait_st();
            if (threadIdx.x == 64) SG_STAMP(l, h, 4);
          } else {
            const float* wo = s_wout + h * 128 + ch * 64;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float hv = fmaxf(__uint_as_float(r0[i]) + bias[i], 0.f);
              acc = fmaf(hv, wo[i], acc);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float hv = fmaxf(__uint_as_float(r1[i]) + bias[32 + i], 0.f);
              acc = fmaf(hv, wo[32 + i], acc);
            }
          }
          tc_fence_before();
          if (p.warp_arrive) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready + h);
          } else {
            mbar_arrive(bar_aready + h);
          }
          if (threadIdx.x == 64) SG_STAMP(l, h, 5);
        }
      }
      if (ch == 1) s_part[t] = acc;
      named_bar(1, 256);
      if (ch == 0 && g < p.total) {
        float v = fmaxf(acc + s_part[t] + s_bout[0], 0.f);     // tf.nn.relu, gfn.py:340
        if (p.use_bbox) {                                     // gfn.py:332-348
          const long long ray = g / p.S;
          const float zz = p.z[g];
          const float px = __fadd_rn(p.rayo[ray * 3 + 0], __fmul_rn(p.rayd[ray * 3 + 0], zz));
          const float py = __fadd_rn(p.rayo[ray * 3 + 1], __fmul_rn(p.rayd[ray * 3 + 1], zz));
          const float pz = __fadd_rn(p.rayo[ray * 3 + 2], __fmul_rn(p.rayd[ray * 3 + 2], zz));
          const bool in = px >= p.bbox[0] && px <= p.bbox[1] && py >= p.bbox[2] &&
                          py <= p.bbox[3] && pz >= p.bbox[4] && pz <= p.bbox[5];
          if (!in) v = 0.f;
        }
        p.sigma[g] = v;
      }
      named_bar(1, 256);     // s_part is free again
    }
  } else if (warp >= 10) {
    // ================================================================ prologue
    const int t = (warp - 10) * 32 + lane;
    for (int it = 0; it < ntile; ++it) {
      const int eb = it & 1;
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      float v[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = 0.f;
      if (g < p.total) {
        const long long ray = g / p.S;
        const float zz = p.z[g];
        float pc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)                          // pts = rayo + rayd * z (gfn.py:264)
          pc[c] = __fadd_rn(p.rayo[ray * 3 + c], __fmul_rn(p.rayd[ray * 3 + c], zz));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          v[c] = pc[c];
          // octaves 0..9 (embedder.py:46-47): accurate sincos at octaves 0, 3, 6, 9 and
          // two double-angle steps after each (keeps the 2^9 x argument error out)
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
        }
      }
      if (it >= 2) mbar_wait(bar_efree + eb, ((it >> 1) - 1) & 1);
      uint8_t* e = s_e + eb * SG_E_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 q;
        q.x = pack2<BF16, 0>(v[8 * j + 0], v[8 * j + 1]);
        q.y = pack2<BF16, 0>(v[8 * j + 2], v[8 * j + 3]);
        q.z = pack2<BF16, 0>(v[8 * j + 4], v[8 * j + 5]);
        q.w = pack2<BF16, 0>(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(e + ((size_t)j * 128 + t) * 16) = q;
        if (ESPLIT) {
          // residual of the 16-bit rounding, again as 16 bit: e = hi + lo to ~2^-22 relative
          float r[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t w = i == 0 ? q.x : (i == 1 ? q.y : (i == 2 ? q.z : q.w));
            r[2 * i] = v[8 * j + 2 * i] - unpack_lo<BF16>(w);
            r[2 * i + 1] = v[8 * j + 2 * i + 1] - unpack_hi<BF16>(w);
          }
          uint4 ql;
          ql.x = pack2<BF16, 0>(r[0], r[1]);
          ql.y = pack2<BF16, 0>(r[2], r[3]);
          ql.z = pack2<BF16, 0>(r[4], r[5]);
          ql.w = pack2<BF16, 0>(r[6], r[7]);
          *reinterpret_cast<uint4*>(s_elo + eb * SG_E_BYTES + ((size_t)j * 128 + t) * 16) = ql;
        }
      }
      fence_proxy_async();
      if (p.warp_arrive) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_eready + eb);
      } else {
        mbar_arrive(bar_eready + eb);
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}


// =====================================================================================
// CTA-pair variant (cta_group::2): the two CTAs of a cluster each own one 128-sample tile
// and HALF of every weight chunk (64 of the 128 output rows of an N-half); the leader CTA
// issues M = 256 MMAs that read both halves, so every SM ingests half the weight bytes.
// Barriers live at identical shared-memory offsets in both CTAs:
//   full[s]      local    producer TMA of this CTA's half landed
//   peerfull[s]  leader   forwarded by the peer CTA once ITS half landed
//   empty[s], dfull[h], efree[b]   both   tcgen05.commit multicast from the leader
//   aready[h], eready[b]           leader one arrival per epilogue / prologue warp of both CTAs
// =====================================================================================
constexpr int S2_NSLOT = 10;
constexpr int S2_SLOT_BYTES = 16384;
constexpr uint32_t S2_LBO = 64 * 16, S2_SBO = 128;
constexpr size_t S2_OFF_RING = 0;
constexpr size_t S2_OFF_E = S2_OFF_RING + (size_t)S2_NSLOT * S2_SLOT_BYTES;
constexpr size_t S2_OFF_BIAS = S2_OFF_E + 2 * SG_E_BYTES;
constexpr size_t S2_OFF_WOUT = S2_OFF_BIAS + 8 * 256 * 4;
constexpr size_t S2_OFF_BOUT = S2_OFF_WOUT + 256 * 4;
constexpr size_t S2_OFF_PART = S2_OFF_BOUT + 16;
constexpr size_t S2_OFF_BAR = S2_OFF_PART + 128 * 4;
// full[10] peerfull[10] empty[10] dfull[2] aready[2] eready[2] efree[2] bar_w = 39
constexpr size_t S2_SMEM = S2_OFF_BAR + 48 * 8;

template <int BF16>
__global__ void __launch_bounds__(SG_THREADS, 1) sigma_tc2_kernel(const SigmaTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_ring = smem + S2_OFF_RING;
  uint8_t* s_e = smem + S2_OFF_E;
  const float* s_bias = reinterpret_cast<const float*>(smem + S2_OFF_BIAS);
  const float* s_wout = reinterpret_cast<const float*>(smem + S2_OFF_WOUT);
  const float* s_bout = reinterpret_cast<const float*>(smem + S2_OFF_BOUT);
  float* s_part = reinterpret_cast<float*>(smem + S2_OFF_PART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S2_OFF_BAR);
  uint64_t* bar_full = bars;             // [10]
  uint64_t* bar_peer = bars + 10;        // [10]
  uint64_t* bar_empty = bars + 20;       // [10]
  uint64_t* bar_dfull = bars + 30;       // [2]
  uint64_t* bar_aready = bars + 32;      // [2]
  uint64_t* bar_eready = bars + 34;      // [2]
  uint64_t* bar_efree = bars + 36;       // [2]
  uint64_t* bar_w = bars + 38;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 40);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int ntile = p.tiles_per_cta;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S2_NSLOT; ++i) {
      mbar_init(bar_full + i, 1); mbar_init(bar_peer + i, 1); mbar_init(bar_empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_dfull + i, 1); mbar_init(bar_aready + i, 16);
      mbar_init(bar_eready + i, 8); mbar_init(bar_efree + i, 1);
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tc2_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t aux_bytes = 8 * 256 * 4 + 256 * 4 + 16;
    mbar_expect_tx(bar_w, aux_bytes);
    bulk_g2s(smem + S2_OFF_BIAS, p.blob + p.off_aux, aux_bytes, bar_w);
  }
  mbar_wait(bar_w, 0);

  if (warp == 1) {
    // ============================================================ TMA producer (own half)
    if (lane == 0) {
      const uint8_t* img = p.blob + p.off_img;
      uint32_t fill = 0;
      for (int it = 0; it < ntile; ++it) {
        uint32_t off = 0;
        for (int l = 0; l < SG_DEPTH; ++l)
          for (int h = 0; h < 2; ++h) {
            const int np = l == 0 ? 1 : (l == SG_SKIP + 1 ? 3 : 2);
            for (int pi = 0; pi < np; ++pi) {
              const int part = l == 0 ? 2 : pi;
              const uint32_t half = part_bytes(part) / 2;
              const uint32_t slot = fill % S2_NSLOT;
              if (fill >= S2_NSLOT) mbar_wait(bar_empty + slot, ((fill / S2_NSLOT) - 1) & 1);
              mbar_expect_tx(bar_full + slot, half);
              bulk_g2s(s_ring + (size_t)slot * S2_SLOT_BYTES, img + off + rank * half, half,
                       bar_full + slot);
              off += 2 * half;
              ++fill;
            }
          }
      }
    }
  } else if (warp == 0 && rank == 1) {
    // ============================================== peer: forward "my half landed" to the leader
    if (lane == 0) {
      uint32_t fill = 0;
      for (int it = 0; it < ntile; ++it)
        for (int c = 0; c < 32; ++c) {
          const uint32_t slot = fill % S2_NSLOT;
          mbar_wait(bar_full + slot, (fill / S2_NSLOT) & 1);
          mbar_arrive_remote(bar_peer + slot, 0);
          ++fill;
        }
    }
  } else if (warp == 0) {
    // ============================================================== MMA issuer (leader CTA)
    if (elect_one()) {
      const uint32_t idesc = make_idesc_mn(BF16, 256, 128);
      const uint32_t ring0 = smem_u32(s_ring), e0 = smem_u32(s_e);
      uint32_t fill = 0, na[2] = {0u, 0u};
      auto wait_a = [&](int h) { mbar_wait(bar_aready + h, na[h] & 1); ++na[h]; };
      for (int it = 0; it < ntile; ++it) {
        const int eb = it & 1;
        mbar_wait(bar_eready + eb, (it >> 1) & 1);
        for (int l = 0; l < SG_DEPTH; ++l) {
          const uint32_t xin = tmem_base + ((l & 1) ? COL_P : COL_Q);
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_t = tmem_base + (h ? COL_D1 : COL_D0);
            const int np = l == 0 ? 1 : (l == SG_SKIP + 1 ? 3 : 2);
            for (int pi = 0; pi < np; ++pi) {
              const int part = l == 0 ? 2 : pi;
              if (l == 0) { if (it > 0) wait_a(h); }
              else if (h == 0 && part < 2) wait_a(part);
              const uint32_t slot = fill % S2_NSLOT;
              mbar_wait(bar_full + slot, (fill / S2_NSLOT) & 1);
              mbar_wait(bar_peer + slot, (fill / S2_NSLOT) & 1);
              tc_fence_after();
              const uint32_t b0 = ring0 + slot * S2_SLOT_BYTES;
              if (part == 2) {
                const uint32_t a0 = e0 + eb * SG_E_BYTES;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc2_mma_ss(d_t, make_b_desc(a0 + ks * 2 * SG_LBO, SG_LBO, SG_SBO),
                             make_b_desc(b0 + ks * 2 * S2_LBO, S2_LBO, S2_SBO), idesc,
                             (l == 0 && ks == 0) ? 0u : 1u);
              } else {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  tc2_mma_ts(d_t, xin + part * 64 + ks * 8,
                             make_b_desc(b0 + ks * 2 * S2_LBO, S2_LBO, S2_SBO), idesc,
                             (part == 0 && ks == 0) ? 0u : 1u);
              }
              tc2_commit_mc(bar_empty + slot, 3);
              ++fill;
            }
            tc2_commit_mc(bar_dfull + h, 3);
          }
          if (l == SG_SKIP + 1) tc2_commit_mc(bar_efree + eb, 3);
        }
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ================================================================ epilogue (both CTAs)
    const int wq = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int t = wq * 32 + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t nd[2] = {0u, 0u};
    for (int it = 0; it < ntile; ++it) {
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      float acc = 0.f;
      for (int l = 0; l < SG_DEPTH; ++l) {
        const uint32_t xout = tb + ((l & 1) ? COL_Q : COL_P);
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_dfull + h, nd[h] & 1);
          ++nd[h];
          tc_fence_after();
          const float* bias = s_bias + l * 256 + h * 128 + ch * 64;
          uint32_t r0[32], r1[32];
          TC_LD32(r0, tb + (h ? COL_D1 : COL_D0) + ch * 64);
          TC_LD32(r1, tb + (h ? COL_D1 : COL_D0) + ch * 64 + 32);
          tc_wait_ld();
          if (l < SG_DEPTH - 1) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 bb = *reinterpret_cast<const float4*>(bias + 4 * i);
              pk[2 * i] = pack2<BF16, 1>(__uint_as_float(r0[4 * i]) + bb.x,
                                             __uint_as_float(r0[4 * i + 1]) + bb.y);
              pk[2 * i + 1] = pack2<BF16, 1>(__uint_as_float(r0[4 * i + 2]) + bb.z,
                                                 __uint_as_float(r0[4 * i + 3]) + bb.w);
            }
            TC_ST16(xout + h * 64 + ch * 32, pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 bb = *reinterpret_cast<const float4*>(bias + 32 + 4 * i);
              pk[2 * i] = pack2<BF16, 1>(__uint_as_float(r1[4 * i]) + bb.x,
                                             __uint_as_float(r1[4 * i + 1]) + bb.y);
              pk[2 * i + 1] = pack2<BF16, 1>(__uint_as_float(r1[4 * i + 2]) + bb.z,
                                                 __uint_as_float(r1[4 * i + 3]) + bb.w);
            }
            TC_ST16(xout + h * 64 + ch * 32 + 16, pk);
            tc_wait_st();
          } else {
            const float* wo = s_wout + h * 128 + ch * 64;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float hv = fmaxf(__uint_as_float(r0[i]) + bias[i], 0.f);
              acc = fmaf(hv, wo[i], acc);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float hv = fmaxf(__uint_as_float(r1[i]) + bias[32 + i], 0.f);
              acc = fmaf(hv, wo[32 + i], acc);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(bar_aready + h);
            else mbar_arrive_remote(bar_aready + h, 0);
          }
        }
      }
      if (ch == 1) s_part[t] = acc;
      named_bar(1, 256);
      if (ch == 0 && g < p.total) {
        float v = fmaxf(acc + s_part[t] + s_bout[0], 0.f);
        if (p.use_bbox) {
          const long long ray = g / p.S;
          const float zz = p.z[g];
          const float px = __fadd_rn(p.rayo[ray * 3 + 0], __fmul_rn(p.rayd[ray * 3 + 0], zz));
          const float py = __fadd_rn(p.rayo[ray * 3 + 1], __fmul_rn(p.rayd[ray * 3 + 1], zz));
          const float pz = __fadd_rn(p.rayo[ray * 3 + 2], __fmul_rn(p.rayd[ray * 3 + 2], zz));
          const bool in = px >= p.bbox[0] && px <= p.bbox[1] && py >= p.bbox[2] &&
                          py <= p.bbox[3] && pz >= p.bbox[4] && pz <= p.bbox[5];
          if (!in) v = 0.f;
        }
        p.sigma[g] = v;
      }
      named_bar(1, 256);
    }
  } else if (warp >= 10) {
    // ================================================================ prologue (both CTAs)
    const int t = (warp - 10) * 32 + lane;
    for (int it = 0; it < ntile; ++it) {
      const int eb = it & 1;
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      float v[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = 0.f;
      if (g < p.total) {
        const long long ray = g / p.S;
        const float zz = p.z[g];
        float pc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          pc[c] = __fadd_rn(p.rayo[ray * 3 + c], __fmul_rn(p.rayd[ray * 3 + c], zz));
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
        }
      }
      if (it >= 2) mbar_wait(bar_efree + eb, ((it >> 1) - 1) & 1);
      uint8_t* e = s_e + eb * SG_E_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 q;
        q.x = pack2<BF16, 0>(v[8 * j + 0], v[8 * j + 1]);
        q.y = pack2<BF16, 0>(v[8 * j + 2], v[8 * j + 3]);
        q.z = pack2<BF16, 0>(v[8 * j + 4], v[8 * j + 5]);
        q.w = pack2<BF16, 0>(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(e + ((size_t)j * 128 + t) * 16) = q;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar_eready + eb);
        else mbar_arrive_remote(bar_eready + eb, 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { __syncwarp(); tc2_dealloc(tmem_base, 512); }
}

template <int BF16>
int launch_sigma2(nf_ctx* ctx, const SigmaTcParams& p, int grid, cudaStream_t st) {
  NF_CUDA(ctx, cudaFuncSetAttribute(sigma_tc2_kernel<BF16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_SMEM));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SG_THREADS);
  cfg.dynamicSmemBytes = S2_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NF_CUDA(ctx, cudaLaunchKernelEx(&cfg, sigma_tc2_kernel<BF16>, p));
  return NF_OK;
}

uint16_t h_bits(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
uint16_t bf_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}

template <int BF16, int CL, int NSLOT, int ESPLIT>
int launch_sigma(nf_ctx* ctx, const SigmaTcParams& p, int grid, cudaStream_t st) {
  constexpr size_t SG_SMEM = SgLayout<NSLOT, ESPLIT>::total;
  NF_CUDA(ctx, cudaFuncSetAttribute(sigma_tc_kernel<BF16, CL, NSLOT, ESPLIT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SG_THREADS);
  cfg.dynamicSmemBytes = SG_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NF_CUDA(ctx, cudaLaunchKernelEx(&cfg, sigma_tc_kernel<BF16, CL, NSLOT, ESPLIT>, p));
  return NF_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// Host packing for the sigma network: append to the blob
//   image (fp16) ++ image (bf16): the weights in the order the kernel streams them:
//     for layer l, for N-half h, for part (l = 0: the 64-row input block; else K-blocks 0, 1;
//     l = skip+1 additionally the 64-row input block): a [kg][128 n][8] K-major,
//     swizzle-free block (kg = 16 k-groups for a K-block, 8 for the input block).
//   aux (fp32): bias[8][256], w_out[256], b_out[4].
int nf_sigma_tc_pack(nf_mlp* m) {
  const nf_mlp_desc& d = m->d;
  if (d.kind != NF_MLP_SIGMA || d.width != 256 || d.depth != 8 || d.skip_at != 4 ||
      d.out_dim != 1 || d.in_dim != 63 || d.n_freqs_a != 10)
    return NF_OK;  // FP32 path only
  const size_t halves = (size_t)(64 + 4 * 256 + (256 + 64) + 2 * 256) * 256;
  const size_t aux_floats = 8 * 256 + 256 + 4;
  size_t base = (m->blob.size() + 255) / 256 * 256;
  m->off_tc_f16 = base;
  m->off_tc_bf16 = base + (halves * 2 + 255) / 256 * 256;
  m->off_tc_aux = m->off_tc_bf16 + (halves * 2 + 255) / 256 * 256;
  m->tc_bytes = halves * 2;
  m->tc_aux_bytes = aux_floats * 4;
  m->blob.resize(m->off_tc_aux + (m->tc_aux_bytes + 255) / 256 * 256, 0);
  uint16_t* i16 = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tc_f16);
  uint16_t* ibf = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tc_bf16);
  float* aux = reinterpret_cast<float*>(m->blob.data() + m->off_tc_aux);
  size_t pos = 0;
  for (int l = 0; l < 8; ++l)
    for (int h = 0; h < 2; ++h) {
      const int np = l == 0 ? 1 : (l == 5 ? 3 : 2);
      for (int pi = 0; pi < np; ++pi) {
        const int part = l == 0 ? 2 : pi;
        const int kk = part == 2 ? 64 : 128;
        const int r0 = part == 2 ? (l == 0 ? 0 : 256) : part * 128;   // first Keras row
        const int kreal = part == 2 ? 63 : 128;
        for (int k = 0; k < kk; ++k)
          for (int n = 0; n < 128; ++n) {
            float v = k < kreal ? d.W[l][(size_t)(r0 + k) * 256 + h * 128 + n] : 0.f;
            size_t idx = pos + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8);
            i16[idx] = h_bits(v);
            ibf[idx] = bf_bits(v);
          }
        pos += (size_t)kk * 128;
      }
    }
  for (int l = 0; l < 8; ++l) memcpy(aux + l * 256, d.b[l], 256 * sizeof(float));
  for (int c = 0; c < 256; ++c) aux[8 * 256 + c] = d.W[8][c];
  aux[8 * 256 + 256] = d.b[8][0];
  // CTA-pair images: every chunk as [rank][kg][64 n][8] (rank r holds output rows 64 r .. 64 r + 63)
  {
    size_t base2 = (m->blob.size() + 255) / 256 * 256;
    m->off_tc2_f16 = base2;
    m->off_tc2_bf16 = base2 + (halves * 2 + 255) / 256 * 256;
    m->blob.resize(m->off_tc2_bf16 + (halves * 2 + 255) / 256 * 256, 0);
    uint16_t* j16 = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tc2_f16);
    uint16_t* jbf = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tc2_bf16);
    size_t pos2 = 0;
    for (int l = 0; l < 8; ++l)
      for (int h = 0; h < 2; ++h) {
        const int np = l == 0 ? 1 : (l == 5 ? 3 : 2);
        for (int pi = 0; pi < np; ++pi) {
          const int part = l == 0 ? 2 : pi;
          const int kk = part == 2 ? 64 : 128;
          const int r0 = part == 2 ? (l == 0 ? 0 : 256) : part * 128;
          const int kreal = part == 2 ? 63 : 128;
          for (int rk = 0; rk < 2; ++rk)
            for (int k = 0; k < kk; ++k)
              for (int n = 0; n < 64; ++n) {
                float v = k < kreal ? d.W[l][(size_t)(r0 + k) * 256 + h * 128 + rk * 64 + n] : 0.f;
                size_t idx = pos2 + (size_t)rk * kk * 64 + ((size_t)(k / 8) * 64 + n) * 8 + (k % 8);
                j16[idx] = h_bits(v);
                jbf[idx] = bf_bits(v);
              }
          pos2 += (size_t)kk * 128;
        }
      }
  }
  return NF_OK;
}

int nf_tc_sigma_launch(nf_ctx* ctx, const nf_mlp* m, const float* rayo, const float* rayd,
                       const float* z, int n_rays, int S, const float* bbox_host, float* sigma,
                       int precision, cudaStream_t st) {
  NF_CHECK_ARG(ctx, m->dev, "network not uploaded (call nf_mlp_upload first)");
  NF_CHECK_ARG(ctx, precision == NF_PREC_F16 || precision == NF_PREC_BF16 ||
                        precision == NF_PREC_F16E, "bad precision");
  const bool esplit = precision == NF_PREC_F16E;
  if (m->tc_bytes == 0)
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "no tcgen05 kernel for this sigma network (need 8 x 256, skip 4, F = 10); "
                        "use NF_PREC_FP32");
  SigmaTcParams p;
  memset(&p, 0, sizeof(p));
  p.blob = (const uint8_t*)m->dev;
  p.off_img = precision == NF_PREC_BF16 ? m->off_tc_bf16 : m->off_tc_f16;
  p.off_aux = m->off_tc_aux;
  p.rayo = rayo; p.rayd = rayd; p.z = z; p.S = S; p.sigma = sigma;
  p.total = (long long)n_rays * S;
  if (bbox_host) { memcpy(p.bbox, bbox_host, sizeof(p.bbox)); p.use_bbox = 1; }
  const long long tiles = (p.total + 127) / 128;
  // Variant selection (tuning knobs, read per call): NF_SIGMA_CLUSTER = 1 | 2 | 4 multicast
  // cluster of the single-CTA kernel (default 2); NF_SIGMA_PAIR = 1 selects the cta_group::2
  // kernel (correct, but ~20 % slower on B200: see DESIGN.md section 6).
  int cl_env = 2, pair_env = 0;
  if (const char* e = getenv("NF_SIGMA_CLUSTER")) {
    cl_env = atoi(e);
    if (cl_env != 1 && cl_env != 2 && cl_env != 4) cl_env = 2;
  }
  if (const char* e = getenv("NF_SIGMA_PAIR")) pair_env = atoi(e);
  if (const char* e = getenv("NF_SIGMA_WARP_ARRIVE")) p.warp_arrive = atoi(e) ? 1 : 0;
  const bool bf = precision == NF_PREC_BF16;
  if (pair_env && !esplit) {
    int grid2 = ctx->sm_count / 2 * 2;
    if (tiles < grid2) grid2 = (int)((tiles + 1) / 2 * 2);
    p.tiles_per_cta = (int)((tiles + grid2 - 1) / grid2);
    p.off_img = bf ? m->off_tc2_bf16 : m->off_tc2_f16;
    return bf ? launch_sigma2<1>(ctx, p, grid2, st) : launch_sigma2<0>(ctx, p, grid2, st);
  }
  int cl = cl_env;
  int grid = ctx->sm_count / cl * cl;
  if (tiles < grid) grid = (int)((tiles + cl - 1) / cl * cl);
  p.tiles_per_cta = (int)((tiles + grid - 1) / grid);
  // NF_PREC_F16E: hi/lo-split positional encoding (4-slot ring); NF_SIGMA_NSLOT=4 runs the
  // plain kernels on the 4-slot ring too (tuning / A-B timing only)
  int nslot = esplit ? 4 : 5;
  if (const char* e = getenv("NF_SIGMA_NSLOT")) { if (atoi(e) == 4) nslot = 4; }
  const char* dbg_path = getenv("NF_SIGMA_DBG");
  if (dbg_path) {
    NF_CUDA(ctx, cudaMalloc(&p.dbg, 96 * sizeof(long long)));
    NF_CUDA(ctx, cudaMemsetAsync(p.dbg, 0, 96 * sizeof(long long), st));
  }
  auto dump_dbg = [&](int rc) {
    if (!dbg_path) return rc;
    long long h[96];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    if (FILE* f = fopen(dbg_path, "w")) {
      for (int i = 0; i < 96; ++i) fprintf(f, "%lld%c", h[i], i % 6 == 5 ? '\n' : ' ');
      fclose(f);
    }
    return rc;
  };
#define NF_SG(B, C)                                                                         \
  (esplit ? launch_sigma<B, C, 4, 1>(ctx, p, grid, st)                                      \
          : (nslot == 4 ? launch_sigma<B, C, 4, 0>(ctx, p, grid, st)                        \
                        : launch_sigma<B, C, 5, 0>(ctx, p, grid, st)))
  if (cl == 1) return dump_dbg(bf ? NF_SG(1, 1) : NF_SG(0, 1));
  if (cl == 2) return dump_dbg(bf ? NF_SG(1, 2) : NF_SG(0, 2));
  return dump_dbg(bf ? NF_SG(1, 4) : NF_SG(0, 4));
#undef NF_SG
}
