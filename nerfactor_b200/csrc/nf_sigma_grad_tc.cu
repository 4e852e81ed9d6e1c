// tcgen05 forward + input-gradient of the NeRF sigma network:
//   sigma = relu(raw(x)),  normal = -l2_normalize(d sigma / d x)
// (the GradientTape.batch_jacobian block of nerfactor/geometry_from_nerf.py:285-300), fp16 or
// bf16 operands with fp32 accumulation.  Same machinery as sigma_tc_kernel (nf_sigma_tc.cu):
// one 128-sample tile per CTA, P/Q activation buffers + D0/D1 accumulator halves in TMEM,
// weights streamed through a shared-memory ring in consumption order -- first the forward
// image, then the backward image (the same W_l, laid out with the INPUT feature as the MMA's
// N index, so G_{l} = (G_{l+1} * relu') W_l^T is again a K-major MMA).
//   * forward epilogues additionally record the ReLU pattern of every layer as bit masks in
//     shared memory; the layer-7 epilogue writes the backward seed w_out * mask_7 into TMEM.
//   * backward epilogues turn an accumulator half into the next A operand by masking.
//   * the two places where the positional encoding enters (skip layer 5, layer 0) produce
//     a [128 x 64] gradient block each (N = 64 MMAs); every epilogue thread immediately
//     contracts its columns with d e_j / d x_c (2^f cos, -2^f sin, identity; embedder.py:46-47)
//     so only three partial sums per thread survive.
#include "nf_common.cuh"
#include "nf_tc_ptx.cuh"
#include <type_traits>

namespace {
using namespace nftc;

constexpr int GT_THREADS = 448;                // MMA, producer, 8 epilogue warps, 4 prologue warps
constexpr int GT_NSLOT = 4;
constexpr int GT_SLOT_BYTES = 32768;
constexpr int GT_E_BYTES = 16384;
constexpr int GT_DEPTH = 8, GT_SKIP = 4;
constexpr int GCOL_P = 0, GCOL_Q = 128, GCOL_D0 = 256, GCOL_D1 = 384;
constexpr uint32_t GT_LBO = 128 * 16, GT_SBO = 128;
constexpr uint32_t GT_LBO64 = 64 * 16;        // N = 64 chunks ([kg][64 n][8])

constexpr size_t GT_OFF_RING = 0;
constexpr size_t GT_OFF_E = GT_OFF_RING + (size_t)GT_NSLOT * GT_SLOT_BYTES;
constexpr size_t GT_OFF_BIAS = GT_OFF_E + 2 * GT_E_BYTES;           // [8][256] f32
constexpr size_t GT_OFF_WOUT = GT_OFF_BIAS + 8 * 256 * 4;            // [256] f32
constexpr size_t GT_OFF_BOUT = GT_OFF_WOUT + 256 * 4;                // [4] f32
constexpr size_t GT_OFF_MASK = GT_OFF_BOUT + 16;                     // [8][8][128] u32
constexpr size_t GT_OFF_PART = GT_OFF_MASK + 8 * 128 * 8 * 4;        // [4][128] f32
constexpr size_t GT_OFF_BAR = GT_OFF_PART + 4 * 128 * 4;
// full[4] empty[4] dfull[2] aready[2] eready[2] efree[2] g gfree bar_w = 19
constexpr size_t GT_SMEM = GT_OFF_BAR + 32 * 8;

struct GradTcParams {
  const uint8_t* blob;
  size_t off_fwd, off_bwd, off_aux;
  const float* rayo;
  const float* rayd;
  const float* z;
  long long total;
  int S;
  int tiles_per_cta;
  float bbox[6];
  int use_bbox;
  float* sigma;      // [n_rays,S]
  float* normal;     // [n_rays,S,3]
};

// forward chunk list of one tile: (layer, half, part) with part 0/1 = hidden K-blocks, 2 = input
// backward chunk list: l = 7, 6: 4 hidden chunks; l = 5: 2 input chunks (N = 64) then 4 hidden;
// l = 4..1: 4 hidden; l = 0: 2 input chunks (N = 64).   Hidden chunk = (half h of the layer's
// INPUT features, K-block kb of its OUTPUT features).

template <int BF16>
__global__ void __launch_bounds__(GT_THREADS, 1) sigma_grad_tc_kernel(const GradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_ring = smem + GT_OFF_RING;
  uint8_t* s_e = smem + GT_OFF_E;
  const float* s_bias = reinterpret_cast<const float*>(smem + GT_OFF_BIAS);
  const float* s_wout = reinterpret_cast<const float*>(smem + GT_OFF_WOUT);
  const float* s_bout = reinterpret_cast<const float*>(smem + GT_OFF_BOUT);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + GT_OFF_MASK);
  float* s_part = reinterpret_cast<float*>(smem + GT_OFF_PART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_OFF_BAR);
  uint64_t* bar_full = bars;            // [4]
  uint64_t* bar_empty = bars + 4;       // [4]
  uint64_t* bar_dfull = bars + 8;       // [2]
  uint64_t* bar_aready = bars + 10;     // [2]
  uint64_t* bar_eready = bars + 12;     // [2]
  uint64_t* bar_efree = bars + 14;      // [2]
  uint64_t* bar_g = bars + 16;          // input-gradient block ready (commit)
  uint64_t* bar_gfree = bars + 17;      // ... consumed (256 arrivals)
  uint64_t* bar_w = bars + 18;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = p.tiles_per_cta;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_NSLOT; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_dfull + i, 1); mbar_init(bar_aready + i, 256);
      mbar_init(bar_eready + i, 128); mbar_init(bar_efree + i, 256);
    }
    mbar_init(bar_g, 1); mbar_init(bar_gfree, 256); mbar_init(bar_w, 1);
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
    const uint32_t aux_bytes = 8 * 256 * 4 + 256 * 4 + 16;
    mbar_expect_tx(bar_w, aux_bytes);
    bulk_g2s(smem + GT_OFF_BIAS, p.blob + p.off_aux, aux_bytes, bar_w);
  }
  mbar_wait(bar_w, 0);

  if (warp == 1) {
    // ============================================================ TMA producer
    if (lane == 0) {
      uint32_t fill = 0;
      auto push = [&](const uint8_t* src, uint32_t bytes) {
        const uint32_t slot = fill % GT_NSLOT;
        if (fill >= GT_NSLOT) mbar_wait(bar_empty + slot, ((fill / GT_NSLOT) - 1) & 1);
        mbar_expect_tx(bar_full + slot, bytes);
        bulk_g2s(s_ring + (size_t)slot * GT_SLOT_BYTES, src, bytes, bar_full + slot);
        ++fill;
      };
      for (int it = 0; it < ntile; ++it) {
        const uint8_t* f = p.blob + p.off_fwd;
        for (int l = 0; l < GT_DEPTH; ++l)
          for (int h = 0; h < 2; ++h) {
            const int np = l == 0 ? 1 : (l == GT_SKIP + 1 ? 3 : 2);
            for (int pi = 0; pi < np; ++pi) {
              const uint32_t bytes = (l == 0 || pi == 2) ? 16384u : 32768u;
              push(f, bytes);
              f += bytes;
            }
          }
        const uint8_t* b = p.blob + p.off_bwd;
        for (int l = GT_DEPTH - 1; l >= 0; --l) {
          if (l == GT_SKIP + 1 || l == 0)
            for (int kb = 0; kb < 2; ++kb) { push(b, 16384u); b += 16384; }
          if (l > 0)
            for (int c = 0; c < 4; ++c) { push(b, 32768u); b += 32768; }
        }
      }
    }
  } else if (warp == 0) {
    // ============================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BF16, 128);
      const uint32_t idesc64 = make_idesc(BF16, 64);
      const uint32_t ring0 = smem_u32(s_ring), e0 = smem_u32(s_e);
      uint32_t fill = 0, na[2] = {0u, 0u}, ngf = 0;
      auto wait_a = [&](int h) { mbar_wait(bar_aready + h, na[h] & 1); ++na[h]; };
      auto slot_wait = [&]() -> uint32_t {
        const uint32_t slot = fill % GT_NSLOT;
        mbar_wait(bar_full + slot, (fill / GT_NSLOT) & 1);
        tc_fence_after();
        return slot;
      };
      auto slot_done = [&](uint32_t slot) { tc_commit(bar_empty + slot); ++fill; };
      for (int it = 0; it < ntile; ++it) {
        const int eb = it & 1;
        mbar_wait(bar_eready + eb, (it >> 1) & 1);
        if (it > 0) { mbar_wait(bar_gfree, ngf & 1); ++ngf; tc_fence_after(); }   // D0 / D1 free
        // ------------------------------------------------------------ forward
        for (int l = 0; l < GT_DEPTH; ++l) {
          const uint32_t xin = tmem_base + ((l & 1) ? GCOL_P : GCOL_Q);
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_t = tmem_base + (h ? GCOL_D1 : GCOL_D0);
            const int np = l == 0 ? 1 : (l == GT_SKIP + 1 ? 3 : 2);
            for (int pi = 0; pi < np; ++pi) {
              const int part = l == 0 ? 2 : pi;
              if (l > 0 && h == 0 && part < 2) wait_a(part);
              const uint32_t slot = slot_wait();
              const uint32_t b0 = ring0 + slot * GT_SLOT_BYTES;
              if (part == 2) {
                const uint32_t a0 = e0 + eb * GT_E_BYTES;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc_mma_ss(d_t, make_b_desc(a0 + ks * 2 * GT_LBO, GT_LBO, GT_SBO),
                            make_b_desc(b0 + ks * 2 * GT_LBO, GT_LBO, GT_SBO), idesc,
                            (l == 0 && ks == 0) ? 0u : 1u);
              } else {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  tc_mma_ts(d_t, xin + part * 64 + ks * 8,
                            make_b_desc(b0 + ks * 2 * GT_LBO, GT_LBO, GT_SBO), idesc,
                            (part == 0 && ks == 0) ? 0u : 1u);
              }
              slot_done(slot);
            }
            tc_commit(bar_dfull + h);
          }
        }
        // ------------------------------------------------------------ backward
        // g_pre of layer l lives in Q for odd l and in P for even l (layer 7's epilogue wrote Q)
        for (int l = GT_DEPTH - 1; l >= 0; --l) {
          const uint32_t ain = tmem_base + ((l & 1) ? GCOL_Q : GCOL_P);
          if (l == GT_SKIP + 1 || l == 0) {
            // input-gradient block: [128 x 64] = g_pre (K = 256) x W_l[input rows]^T -> D0[0:64)
            wait_a(0);
            wait_a(1);
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t slot = slot_wait();
              const uint32_t b0 = ring0 + slot * GT_SLOT_BYTES;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                tc_mma_ts(tmem_base + GCOL_D0, ain + kb * 64 + ks * 8,
                          make_b_desc(b0 + ks * 2 * GT_LBO64, GT_LBO64, GT_SBO), idesc64,
                          (kb == 0 && ks == 0) ? 0u : 1u);
              slot_done(slot);
            }
            tc_commit(bar_g);
            if (l > 0) { mbar_wait(bar_gfree, ngf & 1); ++ngf; tc_fence_after(); }
          }
          if (l == 0) break;
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_t = tmem_base + (h ? GCOL_D1 : GCOL_D0);
            for (int kb = 0; kb < 2; ++kb) {
              if (l != GT_SKIP + 1 && h == 0) wait_a(kb);
              const uint32_t slot = slot_wait();
              const uint32_t b0 = ring0 + slot * GT_SLOT_BYTES;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                tc_mma_ts(d_t, ain + kb * 64 + ks * 8,
                          make_b_desc(b0 + ks * 2 * GT_LBO, GT_LBO, GT_SBO), idesc,
                          (kb == 0 && ks == 0) ? 0u : 1u);
              slot_done(slot);
            }
            tc_commit(bar_dfull + h);
          }
        }
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ================================================================ epilogue
    const int wq = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int t = wq * 32 + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t nd[2] = {0u, 0u}, ng = 0;
    for (int it = 0; it < ntile; ++it) {
      const int eb = it & 1;
      const long long tile = (long long)it * gridDim.x + blockIdx.x;
      const long long g = tile * 128 + t;
      float acc = 0.f;
      // ------------------------------------------------------------ forward
      for (int l = 0; l < GT_DEPTH; ++l) {
        const uint32_t xout = tb + ((l & 1) ? GCOL_Q : GCOL_P);
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_dfull + h, nd[h] & 1);
          ++nd[h];
          tc_fence_after();
          const float* bias = s_bias + l * 256 + h * 128 + ch * 64;
          uint32_t r0[32], r1[32];
          TC_LD32(r0, tb + (h ? GCOL_D1 : GCOL_D0) + ch * 64);
          TC_LD32(r1, tb + (h ? GCOL_D1 : GCOL_D0) + ch * 64 + 32);
          tc_wait_ld();
          uint32_t* mk = s_mask + ((size_t)(l * 8 + h * 4 + ch * 2) * 128 + t);
          const float* wo = s_wout + h * 128 + ch * 64;
          const float gs = s_bout[1];            // power-of-two seed scale (fp16 range)
          // one 32-column group: bias, ReLU pattern, next A operand (or head partial + seed)
          auto group = [&](uint32_t (&r)[32], int gi) {
            const float* bg = bias + 32 * gi;
            uint32_t mm = 0u;
            uint32_t pk[16];
            if (l < GT_DEPTH - 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 bb = *reinterpret_cast<const float4*>(bg + 4 * i);
                const float a0 = __uint_as_float(r[4 * i]) + bb.x, a1 = __uint_as_float(r[4 * i + 1]) + bb.y;
                const float a2 = __uint_as_float(r[4 * i + 2]) + bb.z, a3 = __uint_as_float(r[4 * i + 3]) + bb.w;
                mm |= (a0 > 0.f ? 1u : 0u) << (4 * i) | (a1 > 0.f ? 1u : 0u) << (4 * i + 1) |
                      (a2 > 0.f ? 1u : 0u) << (4 * i + 2) | (a3 > 0.f ? 1u : 0u) << (4 * i + 3);
                pk[2 * i] = pack2<BF16, 1>(a0, a1);
                pk[2 * i + 1] = pack2<BF16, 1>(a2, a3);
              }
              TC_ST16(xout + h * 64 + ch * 32 + 16 * gi, pk);
            } else {
              // head partial + backward seed d raw / d h_8 = w_out, masked, into the Q buffer
              const float* wg = wo + 32 * gi;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float a0 = __uint_as_float(r[2 * i]) + bg[2 * i];
                const float a1 = __uint_as_float(r[2 * i + 1]) + bg[2 * i + 1];
                const float w0 = wg[2 * i], w1 = wg[2 * i + 1];
                acc = fmaf(fmaxf(a0, 0.f), w0, acc);
                acc = fmaf(fmaxf(a1, 0.f), w1, acc);
                mm |= (a0 > 0.f ? 1u : 0u) << (2 * i) | (a1 > 0.f ? 1u : 0u) << (2 * i + 1);
                pk[i] = pack2<BF16, 0>(a0 > 0.f ? gs * w0 : 0.f, a1 > 0.f ? gs * w1 : 0.f);
              }
              TC_ST16(tb + GCOL_Q + h * 64 + ch * 32 + 16 * gi, pk);
            }
            mk[128 * gi] = mm;
          };
          group(r0, 0);
          group(r1, 1);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_aready + h);
        }
      }
      // sigma (tf.nn.relu of the head, gfn.py:291-292) and the sign of raw
      s_part[ch * 128 + t] = acc;
      named_bar(1, 256);
      const float raw = s_part[t] + s_part[128 + t] + s_bout[0];
      const float pos = raw > 0.f ? 1.f : 0.f;
      if (ch == 0 && g < p.total) {
        float v = fmaxf(raw, 0.f);
        if (p.use_bbox) {                                   // gfn.py:275-277, 303-305
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
      // ------------------------------------------------------------ backward
      float gx = 0.f, gy = 0.f, gz = 0.f;
      // contracts this thread's 32 columns of an input-gradient block with d e_j / d x_c
      auto input_block_c = [&](auto CH) {
        constexpr int chc = decltype(CH)::value;
        mbar_wait(bar_g, ng & 1);
        ++ng;
        tc_fence_after();
        uint32_t r[32];
        TC_LD32(r, tb + GCOL_D0 + chc * 32);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(bar_gfree);
        const uint8_t* e = s_e + eb * GT_E_BYTES;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int j = chc * 32 + i;                 // column of the 63-wide encoding (compile time)
          if (j >= 63) continue;
          const float gv = __uint_as_float(r[i]);
          float jac;
          int c;
          if (j < 3) { c = j; jac = 1.f; }
          else {
            const int q = j - 3, f = q / 6, w6 = q % 6;
            c = w6 % 3;
            const int partner = w6 < 3 ? j + 3 : j - 3;  // sin <-> cos of the same (f, c)
            uint16_t hb = *reinterpret_cast<const uint16_t*>(e + ((size_t)(partner >> 3) * 128 + t) * 16 + (partner & 7) * 2);
            float pv;
            if (BF16) pv = __uint_as_float((uint32_t)hb << 16);
            else { __half hh; memcpy(&hh, &hb, 2); pv = __half2float(hh); }
            const float sc = (float)(1 << f);
            jac = w6 < 3 ? sc * pv : -sc * pv;            // d sin = f cos, d cos = -f sin
          }
          if (c == 0) gx = fmaf(gv, jac, gx);
          else if (c == 1) gy = fmaf(gv, jac, gy);
          else gz = fmaf(gv, jac, gz);
        }
      };
      auto input_block = [&]() {
        if (ch == 0) input_block_c(std::integral_constant<int, 0>{});
        else input_block_c(std::integral_constant<int, 1>{});
      };
      for (int l = GT_DEPTH - 1; l >= 1; --l) {
        if (l == GT_SKIP + 1) input_block();
        // G_l half h (inputs h*128 + ch*64 ..) -> mask with relu' of layer l-1 -> A operand
        const uint32_t aout = tb + (((l - 1) & 1) ? GCOL_Q : GCOL_P);
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_dfull + h, nd[h] & 1);
          ++nd[h];
          tc_fence_after();
          uint32_t r0[32], r1[32];
          TC_LD32(r0, tb + (h ? GCOL_D1 : GCOL_D0) + ch * 64);
          TC_LD32(r1, tb + (h ? GCOL_D1 : GCOL_D0) + ch * 64 + 32);
          tc_wait_ld();
          const uint32_t* mk = s_mask + ((size_t)((l - 1) * 8 + h * 4 + ch * 2) * 128 + t);
          const uint32_t m0 = mk[0], m1 = mk[128];
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack2<BF16, 0>((m0 >> (2 * i)) & 1u ? __uint_as_float(r0[2 * i]) : 0.f,
                                   (m0 >> (2 * i + 1)) & 1u ? __uint_as_float(r0[2 * i + 1]) : 0.f);
          TC_ST16(aout + h * 64 + ch * 32, pk);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack2<BF16, 0>((m1 >> (2 * i)) & 1u ? __uint_as_float(r1[2 * i]) : 0.f,
                                   (m1 >> (2 * i + 1)) & 1u ? __uint_as_float(r1[2 * i + 1]) : 0.f);
          TC_ST16(aout + h * 64 + ch * 32 + 16, pk);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_aready + h);
        }
      }
      input_block();                                   // layer 0
      mbar_arrive(bar_efree + eb);                     // last reader of this tile's encoding
      // combine the two column halves, apply relu'(raw), normalise (gfn.py:293-297)
      if (ch == 1) { s_part[128 + t] = gx; s_part[256 + t] = gy; s_part[384 + t] = gz; }
      named_bar(1, 256);
      if (ch == 0 && g < p.total) {
        const float un = pos * s_bout[2];                 // undo the seed scale
        const float ax = (gx + s_part[128 + t]) * un;
        const float ay = (gy + s_part[256 + t]) * un;
        const float az = (gz + s_part[384 + t]) * un;
        const float s = 1.0f / sqrtf(fmaxf(ax * ax + ay * ay + az * az, 1e-12f));
        p.normal[g * 3 + 0] = -ax * s;
        p.normal[g * 3 + 1] = -ay * s;
        p.normal[g * 3 + 2] = -az * s;
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
      uint8_t* e = s_e + eb * GT_E_BYTES;
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

uint16_t gh_bits(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
uint16_t gbf_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}

}  // namespace

// Backward weight image of a sigma network (appended after nf_sigma_tc_pack's images):
// chunks in the order the kernel streams them (see sigma_grad_tc_kernel), fp16 then bf16.
//   hidden chunk (l, h, kb):  [16 kg][128 n][8], element (n, k) = W_l[h*128 + n][kb*128 + k]
//   input  chunk (l, kb):     [16 kg][ 64 n][8], element (n, k) = W_l[r0 + n][kb*128 + k],
//                             r0 = 256 for the skip layer, 0 for layer 0; row 63 is zero.
int nf_sigma_grad_tc_pack(nf_mlp* m) {
  const nf_mlp_desc& d = m->d;
  if (d.kind != NF_MLP_SIGMA || m->tc_bytes == 0) return NF_OK;   // needs the forward images too
  const size_t halves = (size_t)7 * 4 * 128 * 128 + (size_t)4 * 64 * 128;
  size_t base = (m->blob.size() + 255) / 256 * 256;
  m->off_tcb_f16 = base;
  m->off_tcb_bf16 = base + (halves * 2 + 255) / 256 * 256;
  m->blob.resize(m->off_tcb_bf16 + (halves * 2 + 255) / 256 * 256, 0);
  uint16_t* i16 = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tcb_f16);
  uint16_t* ibf = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tcb_bf16);
  {
    // seed scale: power of two bringing max |w_out| into [1, 2) -- the backward pass is linear
    // in it, fp16 keeps its relative precision and the result is unscaled in fp32
    float* aux = reinterpret_cast<float*>(m->blob.data() + m->off_tc_aux);
    float mx = 0.f;
    for (int c = 0; c < 256; ++c) mx = fmaxf(mx, fabsf(d.W[8][c]));
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)
    aux[8 * 256 + 256 + 1] = ldexpf(1.f, 1 - e);
    aux[8 * 256 + 256 + 2] = ldexpf(1.f, e - 1);
  }
  size_t pos = 0;
  for (int l = 7; l >= 0; --l) {
    if (l == 5 || l == 0) {
      const int r0 = l == 5 ? 256 : 0;
      for (int kb = 0; kb < 2; ++kb) {
        for (int n = 0; n < 64; ++n)
          for (int k = 0; k < 128; ++k) {
            float v = n < 63 ? d.W[l][(size_t)(r0 + n) * 256 + kb * 128 + k] : 0.f;
            size_t idx = pos + ((size_t)(k / 8) * 64 + n) * 8 + (k % 8);
            i16[idx] = gh_bits(v);
            ibf[idx] = gbf_bits(v);
          }
        pos += (size_t)64 * 128;
      }
    }
    if (l > 0) {
      for (int h = 0; h < 2; ++h)
        for (int kb = 0; kb < 2; ++kb) {
          for (int n = 0; n < 128; ++n)
            for (int k = 0; k < 128; ++k) {
              float v = d.W[l][(size_t)(h * 128 + n) * 256 + kb * 128 + k];
              size_t idx = pos + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8);
              i16[idx] = gh_bits(v);
              ibf[idx] = gbf_bits(v);
            }
          pos += (size_t)128 * 128;
        }
    }
  }
  return NF_OK;
}

int nf_tc_sigma_grad_launch(nf_ctx* ctx, const nf_mlp* m, const float* rayo, const float* rayd,
                            const float* z, int n_rays, int S, const float* bbox_host,
                            float* sigma, float* normal, int precision, cudaStream_t st) {
  NF_CHECK_ARG(ctx, m->dev, "network not uploaded (call nf_mlp_upload first)");
  NF_CHECK_ARG(ctx, precision == NF_PREC_F16 || precision == NF_PREC_BF16, "bad precision");
  if (m->off_tcb_f16 == 0)
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "no tcgen05 gradient kernel for this sigma network (need 8 x 256, skip 4, "
                        "F = 10); use NF_PREC_FP32");
  GradTcParams p;
  memset(&p, 0, sizeof(p));
  p.blob = (const uint8_t*)m->dev;
  const bool bf = precision == NF_PREC_BF16;
  p.off_fwd = bf ? m->off_tc_bf16 : m->off_tc_f16;
  p.off_bwd = bf ? m->off_tcb_bf16 : m->off_tcb_f16;
  p.off_aux = m->off_tc_aux;
  p.rayo = rayo; p.rayd = rayd; p.z = z; p.S = S; p.sigma = sigma; p.normal = normal;
  p.total = (long long)n_rays * S;
  if (bbox_host) { memcpy(p.bbox, bbox_host, sizeof(p.bbox)); p.use_bbox = 1; }
  const long long tiles = (p.total + 127) / 128;
  int grid = ctx->sm_count;
  if (tiles < grid) grid = (int)tiles;
  p.tiles_per_cta = (int)((tiles + grid - 1) / grid);
  if (bf) {
    NF_CUDA(ctx, cudaFuncSetAttribute(sigma_grad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    sigma_grad_tc_kernel<1><<<grid, GT_THREADS, GT_SMEM, st>>>(p);
  } else {
    NF_CUDA(ctx, cudaFuncSetAttribute(sigma_grad_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    sigma_grad_tc_kernel<0><<<grid, GT_THREADS, GT_SMEM, st>>>(p);
  }
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
