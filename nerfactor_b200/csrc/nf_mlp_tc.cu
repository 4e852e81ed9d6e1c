// tcgen05 fused skip-MLP forward for the per-(point, light) networks
// (light visibility: nerfactor/models/shape.py:213-237; learned BRDF:
// nerfactor/models/nerfactor.py:413-457), sm_100a only.
//
// Design (DESIGN.md "K2"):
//  * persistent CTA per SM, 12 warps: warp 0 = MMA issuer (+TMEM owner), warps
//    4-7 and 8-11 = two worker groups, each owning one 128-row tile at a time
//    (one tile = 128 consecutive light directions of ONE surface point).
//  * the Dense chain never leaves the SM: accumulators D[128x128] fp32 live in
//    TMEM; the bias+ReLU epilogue reads D with tcgen05.ld, packs to fp16/bf16 and
//    writes the next layer's A operand back to TMEM with tcgen05.st; the MMA
//    reads A from TMEM and the weights (B) from shared memory, where the whole
//    network is resident (loaded once per CTA with cp.async.bulk).
//  * the xyz part of the input (63 of 90 columns, identical for all rows of a
//    tile) is folded into a per-point fp32 bias: beff = b + W[:63]^T embed(xyz),
//    computed on the CUDA cores once per point; the tensor cores only contract
//    the per-light columns (27 -> K = 32) and the hidden layers.  Same maths as
//    x @ W + b, fewer MMA flops, and the 2^9-frequency features stay in fp32.
//  * the two groups ping-pong: while one group runs its epilogue the other
//    group's MMAs occupy the tensor pipe.
#include "nf_common.cuh"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nftc;

__device__ __forceinline__ void group_bar256(int id) {  // 256-thread named barrier
  asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory");
}

constexpr int TC_WIDTH = 128;
constexpr int TC_THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr int GRP_COLS = 256;      // TMEM columns per worker group
constexpr int COL_D = 0;           // D accumulator: 128 fp32 columns
constexpr int COL_AH = 128;        // hidden activations: 64 columns (128 x 16-bit)
constexpr int COL_AE = 192;        // per-row embedding: <= 16 columns (32 x 16-bit)

// ------------------------------------------------------------- kernel params
struct TcParams {
  const uint8_t* blob;   // packed network (device)
  size_t off_img;        // B-operand images (fp16 or bf16), see nf_tc_pack
  size_t off_aux;        // fp32 side block
  int kind;              // NF_MLP_LVIS | NF_MLP_BRDF
  int n;                 // surface points
  int L;                 // light directions
  int nr;                // per-point input columns folded into the bias (63 | z_dim)
  int n_freqs_a, n_freqs_b, z_dim;
  int out_act;
  float xyz_scale;
  const float* xyz;      // [n,3]
  const float* xyz_dir;  // [n,3] or NULL: origin of the light directions when it differs from xyz
                         // (jittered evaluation: network at xyz + noise, directions of xyz; shape.py:170)
  const float* lxyz;     // [L,3]
  const float* normal;   // BRDF: [n,3]
  const float* cam;      // BRDF: [n,3]
  const float* zlat;     // BRDF: [n,z_dim]
  float* out;            // [n, L]  (NULL: the light-visibility values are not materialised)
  // ---- fused rendering equation (nf_stageB_fused_fwd: microfacet BRDF, one env-map, L <= 512)
  const float* f_normal;   // [n,3] predicted normals
  const float* f_cam;      // [n,3] camera position per point
  const float* f_albedo;   // [n,3]
  const float* f_rough;    // [n]
  const float* f_lareas;   // [L]
  const float* f_light;    // [P,3] env-map texels (clipped >= 0)
  const int* f_light_idx;  // [L] light -> texel, or NULL
  float f_f0;
  int f_srgb;
  float* f_rgb;            // [n,3]; NULL = plain network evaluation
  const float* cull_normal;  // visibility network, CULL = 1: [n,3] shading normals -- only the
                             // lights with cos(normal, light) > -1e-5 are evaluated, the others get 0
};

// fp32 side block layout (floats): see nf_tc_pack
constexpr int AUX_B = 0;                    // b0..b3: 4 x 128
constexpr int AUX_WOUT = 4 * 128;           // w_out: 128
constexpr int AUX_BOUT = 5 * 128;           // b_out: 1 (padded to 4)
constexpr int AUX_WX0 = 5 * 128 + 4;        // W0 rows [0, nr): nr_pad x 128
// followed by Wx3: W3 rows [128, 128 + nr): nr_pad x 128

template <int KIND>
struct KindCfg;
template <>
struct KindCfg<NF_MLP_LVIS> {
  static constexpr int KE = 32;       // per-row embedding K (27 padded)
  static constexpr int NR_PAD = 64;   // per-point columns (63 padded)
};
template <>
struct KindCfg<NF_MLP_BRDF> {
  static constexpr int KE = 16;       // 15 padded
  static constexpr int NR_PAD = 8;    // z_dim <= 8
};

template <int KIND>
constexpr int img_halves() {
  return (KindCfg<KIND>::KE + 128 + 128 + 128 + KindCfg<KIND>::KE) * TC_WIDTH;
}

template <int KIND>
struct SmemLayout {
  static constexpr int KE = KindCfg<KIND>::KE;
  static constexpr int NR_PAD = KindCfg<KIND>::NR_PAD;
  static constexpr size_t img_bytes = (size_t)img_halves<KIND>() * 2;
  static constexpr size_t aux_floats = AUX_WX0 + 2 * (size_t)NR_PAD * 128;
  static constexpr size_t off_img = 0;
  static constexpr size_t off_aux = img_bytes;
  static constexpr size_t off_beff = off_aux + aux_floats * 4;          // [2 groups][2][128]
  static constexpr size_t off_e = off_beff + 2 * 2 * 128 * 4;           // [2 groups][64]
  static constexpr size_t off_lx = off_e + 2 * 64 * 4;                  // float4 [Lmax]
  static constexpr int LMAX = 1024;
  static constexpr size_t off_bar = off_lx + (size_t)LMAX * 16;
  static constexpr size_t total = off_bar + 128;
};

template <int KIND, int BF16>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const TcParams p) {
  using SL = SmemLayout<KIND>;
  constexpr int KE = SL::KE;
  constexpr int NR_PAD = SL::NR_PAD;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_img = smem + SL::off_img;
  float* s_aux = reinterpret_cast<float*>(smem + SL::off_aux);
  float* s_beff = reinterpret_cast<float*>(smem + SL::off_beff);
  float* s_e = reinterpret_cast<float*>(smem + SL::off_e);
  float4* s_lx = reinterpret_cast<float4*>(smem + SL::off_lx);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL::off_bar);
  uint64_t* bar_w = bars + 0;          // weights landed
  uint64_t* bar_a = bars + 1;          // [2] A operand ready (128 arrivals)
  uint64_t* bar_d = bars + 3;          // [2] D accumulator ready (tcgen05.commit)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = (p.L + 127) / 128;

  // ---------------------------------------------------------------- set-up
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a + 0, 128); mbar_init(bar_a + 1, 128);
    mbar_init(bar_d + 0, 1); mbar_init(bar_d + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int l = threadIdx.x; l < p.L; l += blockDim.x)
    s_lx[l] = make_float4(p.lxyz[l * 3], p.lxyz[l * 3 + 1], p.lxyz[l * 3 + 2], 0.f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t aux_bytes = (uint32_t)(SL::aux_floats * 4);
    mbar_expect_tx(bar_w, (uint32_t)SL::img_bytes + aux_bytes);
    // TMA bulk copies (UBLKCP): whole network -> shared memory, once per CTA
    const uint8_t* gi = p.blob + p.off_img;
    for (size_t o = 0; o < SL::img_bytes; o += 32768) {
      size_t nb = SL::img_bytes - o < 32768 ? SL::img_bytes - o : 32768;
      bulk_g2s(s_img + o, gi + o, (uint32_t)nb, bar_w);
    }
    const uint8_t* ga = p.blob + p.off_aux;
    for (size_t o = 0; o < aux_bytes; o += 32768) {
      size_t nb = aux_bytes - o < 32768 ? aux_bytes - o : 32768;
      bulk_g2s(reinterpret_cast<uint8_t*>(s_aux) + o, ga + o, (uint32_t)nb, bar_w);
    }
  }
  mbar_wait(bar_w, 0);

  // unit of work = one surface point; group G takes points G, G + 2*grid, ...
  const int n_groups = gridDim.x * 2;
  auto group_points = [&](int g) {
    int G = blockIdx.x * 2 + g;
    return G < p.n ? (p.n - 1 - G) / n_groups + 1 : 0;
  };

  if (warp == 0) {
    // =============================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BF16, TC_WIDTH);
      const uint32_t lbo = TC_WIDTH * 16, sbo = 128;
      const uint32_t img0 = smem_u32(s_img);
      // segment byte offsets inside the image: [KE | 128 | 128 | 128 | KE] x 128 x 2 B
      const uint32_t seg_off[5] = {0u, (uint32_t)KE * 256u, (uint32_t)(KE + 128) * 256u,
                                   (uint32_t)(KE + 256) * 256u, (uint32_t)(KE + 384) * 256u};
      const int nt0 = group_points(0) * chunks, nt1 = group_points(1) * chunks;
      const int nt_max = nt0 > nt1 ? nt0 : nt1;
      uint32_t ph[2] = {0u, 0u};
      for (int it = 0; it < nt_max; ++it) {
        for (int layer = 0; layer < 4; ++layer) {
          for (int g = 0; g < 2; ++g) {
            if (it >= (g == 0 ? nt0 : nt1)) continue;
            mbar_wait(bar_a + g, ph[g]);
            ph[g] ^= 1u;
            tc_fence_after();
            const uint32_t tb = tmem_base + g * GRP_COLS;
            const uint32_t d_t = tb + COL_D;
            if (layer == 0) {
#pragma unroll
              for (int k = 0; k < KE / 16; ++k)
                tc_mma_ts(d_t, tb + COL_AE + k * 8,
                          make_b_desc(img0 + seg_off[0] + k * 2 * lbo, lbo, sbo), idesc, k > 0);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                tc_mma_ts(d_t, tb + COL_AH + k * 8,
                          make_b_desc(img0 + seg_off[layer] + k * 2 * lbo, lbo, sbo), idesc, k > 0);
              if (layer == 3) {
#pragma unroll
                for (int k = 0; k < KE / 16; ++k)
                  tc_mma_ts(d_t, tb + COL_AE + k * 8,
                            make_b_desc(img0 + seg_off[4] + k * 2 * lbo, lbo, sbo), idesc, 1u);
              }
            }
            tc_commit(bar_d + g);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================== workers
    const int g = (warp - 4) >> 2;            // group 0 | 1
    const int wq = warp & 3;                  // TMEM lane quarter this warp may access
    const int t = wq * 32 + lane;             // row of the tile == TMEM lane
    const int tg = t;                         // thread index inside the group
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t tb = tmem_base + g * GRP_COLS + lane_addr;
    float* beff0 = s_beff + g * 256;
    float* beff3 = beff0 + 128;
    float* e_s = s_e + g * 64;
    const float* Wx0 = s_aux + AUX_WX0;
    const float* Wx3 = Wx0 + NR_PAD * 128;
    const int G = blockIdx.x * 2 + g;
    uint32_t phd = 0u;

    for (int pt = G; pt < p.n; pt += n_groups) {
      // ---------------------------------------------- per-point (once per L lights)
      const f3 x = ld3(p.xyz + (size_t)pt * 3);
      f3 fr_t, fr_b, fr_n, v_loc;
      if (KIND == NF_MLP_LVIS) {
        // embed(xyz_scale * xyz): embedder.py:46-47
        if (tg < 3) e_s[tg] = (tg == 0 ? x.x : (tg == 1 ? x.y : x.z)) * p.xyz_scale;
        else if (tg < 3 + 3 * p.n_freqs_a) {
          int idx = tg - 3, f = idx / 3, c = idx % 3;
          float xv = (c == 0 ? x.x : (c == 1 ? x.y : x.z)) * p.xyz_scale;
          float s, co;
          sincosf(xv * (float)(1 << f), &s, &co);
          e_s[3 + 6 * f + c] = s;
          e_s[3 + 6 * f + 3 + c] = co;
        }
      } else {
        if (tg < p.z_dim) e_s[tg] = p.zlat[(size_t)pt * p.z_dim + tg];
        world2local_dev(ld3(p.normal + (size_t)pt * 3), fr_t, fr_b, fr_n);   // geom.py:119-149
        f3 v = l2n(ld3(p.cam + (size_t)pt * 3) - x, 1e-6f);                   // shape.py:137-144
        v_loc = mk3(dot3(fr_t, v), dot3(fr_b, v), dot3(fr_n, v));             // nerfactor.py:418
      }
      group_bar(1 + g);
      {
        float a0 = s_aux[AUX_B + 0 * 128 + tg], a3 = s_aux[AUX_B + 3 * 128 + tg];
        for (int k = 0; k < p.nr; ++k) {
          float ev = e_s[k];
          a0 = fmaf(ev, Wx0[k * 128 + tg], a0);
          a3 = fmaf(ev, Wx3[k * 128 + tg], a3);
        }
        beff0[tg] = a0;
        beff3[tg] = a3;
      }
      group_bar(1 + g);

      for (int c = 0; c < chunks; ++c) {
        const int li = c * 128 + t;
        const int lc = li < p.L ? li : p.L - 1;
        // ------------------------------------------------ per-row embedding -> A_e
        float mask = 1.f;
        {
          float4 lp = s_lx[lc];
          f3 d = l2n(mk3(lp.x, lp.y, lp.z) - x, 1e-6f);                       // shape.py:128-135
          float v[KE];
#pragma unroll
          for (int i = 0; i < KE; ++i) v[i] = 0.f;
          f3 q;
          int nf;
          if (KIND == NF_MLP_LVIS) { q = d; nf = 4; }
          else {
            f3 l_loc = mk3(dot3(fr_t, d), dot3(fr_b, d), dot3(fr_n, d));      // nerfactor.py:419
            mask = l_loc.z > 0.f ? 1.f : 0.f;                                 // :429-432
            q = dir2rusink_dev(l_loc, v_loc);                                 // geom.py:152-192
            nf = 2;
          }
          v[0] = q.x; v[1] = q.y; v[2] = q.z;
          float sx, cx, sy, cy, sz, cz;
          sincosf(q.x, &sx, &cx); sincosf(q.y, &sy, &cy); sincosf(q.z, &sz, &cz);
#pragma unroll
          for (int f = 0; f < (KIND == NF_MLP_LVIS ? 4 : 2); ++f) {
            if (f < nf) {
              v[3 + 6 * f + 0] = sx; v[3 + 6 * f + 1] = sy; v[3 + 6 * f + 2] = sz;
              v[3 + 6 * f + 3] = cx; v[3 + 6 * f + 4] = cy; v[3 + 6 * f + 5] = cz;
              // double-angle step to the next octave
              float nsx = 2.f * sx * cx, ncx = 1.f - 2.f * sx * sx;
              float nsy = 2.f * sy * cy, ncy = 1.f - 2.f * sy * sy;
              float nsz = 2.f * sz * cz, ncz = 1.f - 2.f * sz * sz;
              sx = nsx; cx = ncx; sy = nsy; cy = ncy; sz = nsz; cz = ncz;
            }
          }
          uint32_t pk[KE / 2];
#pragma unroll
          for (int i = 0; i < KE / 2; ++i) pk[i] = pack2<BF16, 0>(v[2 * i], v[2 * i + 1]);
          if (KE == 32) { TC_ST16(tb + COL_AE, pk); }
          else { TC_ST8(tb + COL_AE, pk); }
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_a + g);

        // ------------------------------------------------ layers 0..2: epilogue -> A_h
        for (int layer = 0; layer < 3; ++layer) {
          const float* bias = layer == 0 ? beff0 : (s_aux + AUX_B + layer * 128);
          mbar_wait(bar_d + g, phd);
          phd ^= 1u;
          tc_fence_after();
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t r0[32], r1[32];
            TC_LD32(r0, tb + COL_D + c2 * 64);
            TC_LD32(r1, tb + COL_D + c2 * 64 + 32);
            tc_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 bb = *reinterpret_cast<const float4*>(bias + c2 * 64 + 4 * i);
              pk[2 * i] = pack2<BF16, 1>(__uint_as_float(r0[4 * i]) + bb.x,
                                             __uint_as_float(r0[4 * i + 1]) + bb.y);
              pk[2 * i + 1] = pack2<BF16, 1>(__uint_as_float(r0[4 * i + 2]) + bb.z,
                                                 __uint_as_float(r0[4 * i + 3]) + bb.w);
            }
            TC_ST16(tb + COL_AH + c2 * 32, pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 bb = *reinterpret_cast<const float4*>(bias + c2 * 64 + 32 + 4 * i);
              pk[2 * i] = pack2<BF16, 1>(__uint_as_float(r1[4 * i]) + bb.x,
                                             __uint_as_float(r1[4 * i + 1]) + bb.y);
              pk[2 * i + 1] = pack2<BF16, 1>(__uint_as_float(r1[4 * i + 2]) + bb.z,
                                                 __uint_as_float(r1[4 * i + 3]) + bb.w);
            }
            TC_ST16(tb + COL_AH + c2 * 32 + 16, pk);
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_a + g);
        }
        // ------------------------------------------------ layer 3 + head
        mbar_wait(bar_d + g, phd);
        phd ^= 1u;
        tc_fence_after();
        float acc = 0.f;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t r0[32], r1[32];
          TC_LD32(r0, tb + COL_D + c2 * 64);
          TC_LD32(r1, tb + COL_D + c2 * 64 + 32);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float h = fmaxf(__uint_as_float(r0[i]) + beff3[c2 * 64 + i], 0.f);
            acc = fmaf(h, s_aux[AUX_WOUT + c2 * 64 + i], acc);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float h = fmaxf(__uint_as_float(r1[i]) + beff3[c2 * 64 + 32 + i], 0.f);
            acc = fmaf(h, s_aux[AUX_WOUT + c2 * 64 + 32 + i], acc);
          }
        }
        float o = acc + s_aux[AUX_BOUT];
        o = apply_act(p.out_act, o) * mask;
        if (li < p.L) p.out[(size_t)pt * p.L + li] = o;
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// Version 2 (default): the bias add moves INTO the tensor-core contraction.
//
// Measured on B200 (profiles/r2_k2_analysis.md): in the kernel above the tensor pipe is only 45 %
// active because a layer's epilogue (TMEM load -> +bias -> ReLU -> fp16 -> TMEM store, ~230
// instructions per thread plus two exposed tcgen05.ld round trips of ~200 clocks each) is much
// longer than the other group's 512-clock MMA phase it should hide behind; more epilogue warps or
// a readiness-ordered issuer make it slower (issue contention with the MMA thread).  Here
//   * every layer gets one extra K = 16 MMA block whose A operand is the constant row
//     (1, 1, 0, ..) and whose B operand holds the bias as an fp16 hi + lo pair (b = hi + lo to
//     2^-22): D = A W + 1 * b_hi + 1 * b_lo comes out of the tensor core with the bias already
//     added in fp32.  The static biases of layers 1, 2 are part of the weight image; the folded
//     per-point biases of layer 0 and of the skip layer are written by the group into its own
//     4 KB shared-memory block once per surface point.
//   * the epilogue is then just load -> cvt.relu.f16x2 -> store (~80 instructions per thread),
//     with the second half of the accumulator in flight while the first half is converted.
// One extra 128 x 128 x 16 MMA per layer costs 64 of ~512 clocks; the tensor pipe stays busy.
//   * software pipelining across the tiles of a point: the per-row embedding of tile c + 1 is
//     written (second A_e buffer) while layer 3 of tile c is in the tensor pipe, and the head of
//     tile c runs from registers AFTER layer 0 of tile c + 1 has been handed over.
template <int KIND>
struct SmemLayout2 {
  static constexpr int KE = KindCfg<KIND>::KE;
  static constexpr int NR_PAD = KindCfg<KIND>::NR_PAD;
  // image: [W0e KE | W1 128 | W2 128 | W3h 128 | W3e KE | bias1 16 | bias2 16] x 128 x 2 B
  static constexpr size_t img_bytes = ((size_t)img_halves<KIND>() + 2 * 16 * 128) * 2;
  static constexpr size_t aux_floats = AUX_WX0 + 2 * (size_t)NR_PAD * 128;
  static constexpr size_t off_img = 0;
  static constexpr size_t off_bdyn = img_bytes;                          // [2 groups][2][16 x 128] 16-bit
  static constexpr size_t off_aux = off_bdyn + 2 * 2 * 4096;
  static constexpr size_t off_e = off_aux + aux_floats * 4;              // [2 groups][64] f32
  static constexpr size_t off_lx = off_e + 2 * 64 * 4;                   // float4 [Lmax]
  static constexpr int LMAX = 1024;
  static constexpr size_t off_bar = off_lx + (size_t)LMAX * 16;
  static constexpr size_t off_red = off_bar + 128;                       // [2 groups][16][4] f32
  static constexpr size_t off_cnt = off_red + 2 * 16 * 4 * 4;            // [2 groups][8] int: warp counts, tiles
  static constexpr size_t off_list = off_cnt + 2 * 8 * 4;                // [2 groups][4 warps][256] u16
  static constexpr size_t total = off_list + 2 * 4 * 256 * 2;
};
constexpr int COL_ONE = 208;       // constant A operand (1, 1, 0, ...): 8 columns
constexpr int COL_AE1 = 216;       // second per-row embedding buffer (tiles alternate: the next
                                   // tile's embedding is written while layer 3 still reads this one)

// SELF = 1 (learned-BRDF network): no separate MMA-issuer warp -- thread 0 of a worker group issues
// its own group's MMAs after the group's named barrier, because the number of tiles per point is
// data-dependent there (front-lit compaction).  SELF = 0 (visibility network): warp 0 issues for both
// groups in strict alternation (measured faster: 32.4 vs 40.9 ms).  The issuer does not know how
// many tiles a group will run: it serves hand-overs (layer 0, 1, 2, 3, 0, ... of running tile
// k = 0, 1, ...) until the group raises its `done` flag with a last arrival.
// CULL = 1: only the front-lit lights of a point become tile rows (always for the BRDF network,
// nerfactor.py:429-458; for the visibility network when the caller only needs the rendered colour:
// nerfactor.py:329-330 multiplies the visibility of every other light by zero).
template <int KIND, int BF16, int SELF, int CULL>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc2_kernel(const TcParams p) {
  static_assert(KIND != NF_MLP_BRDF || CULL == 1, "the BRDF network always runs on the front-lit lights");
  using SL = SmemLayout2<KIND>;
  constexpr int KE = SL::KE;
  constexpr int NR_PAD = SL::NR_PAD;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_img = smem + SL::off_img;
  uint8_t* s_bdyn = smem + SL::off_bdyn;
  float* s_aux = reinterpret_cast<float*>(smem + SL::off_aux);
  float* s_e = reinterpret_cast<float*>(smem + SL::off_e);
  float4* s_lx = reinterpret_cast<float4*>(smem + SL::off_lx);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL::off_bar);
  uint64_t* bar_w = bars + 0;          // weights landed
  uint64_t* bar_a = bars + 1;          // [2] A operand (first K-half) ready (128 arrivals)
  uint64_t* bar_d = bars + 3;          // [2] D accumulator ready (tcgen05.commit)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 5);
  float* s_red = reinterpret_cast<float*>(smem + SL::off_red);
  volatile int* s_cnt = reinterpret_cast<volatile int*>(smem + SL::off_cnt);   // [g][0..3] front-lit lights per warp, [g][4] tiles of the point
  uint16_t* s_list = reinterpret_cast<uint16_t*>(smem + SL::off_list);         // BRDF: compacted front-lit light indices
  // fused rendering equation: texel * area per light in the upper half of the light table
  const bool fuse = KIND == NF_MLP_LVIS && p.f_rgb != nullptr;
  float4* s_lrgb = s_lx + 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = (p.L + 127) / 128;

  // ---------------------------------------------------------------- set-up
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a + 0, 128); mbar_init(bar_a + 1, 128);
    mbar_init(bar_d + 0, 1); mbar_init(bar_d + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int l = threadIdx.x; l < p.L; l += blockDim.x) {
    s_lx[l] = make_float4(p.lxyz[l * 3], p.lxyz[l * 3 + 1], p.lxyz[l * 3 + 2], 0.f);
    if (fuse) {            // light * area: rgb += brdf (lvis light) cos area, nerfactor.py:333-336
      const float* t = p.f_light + (size_t)(p.f_light_idx ? p.f_light_idx[l] : l) * 3;
      const float ar = p.f_lareas[l];
      s_lrgb[l] = make_float4(t[0] * ar, t[1] * ar, t[2] * ar, 0.f);
    }
  }
  for (int i = threadIdx.x; i < 2 * 2 * 4096 / 4; i += blockDim.x)       // k = 2..15 rows stay zero
    reinterpret_cast<uint32_t*>(s_bdyn)[i] = 0u;
  if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0;                            // incl. the `done` flags [g][5]
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t aux_bytes = (uint32_t)(SL::aux_floats * 4);
    mbar_expect_tx(bar_w, (uint32_t)SL::img_bytes + aux_bytes);
    const uint8_t* gi = p.blob + p.off_img;
    for (size_t o = 0; o < SL::img_bytes; o += 32768) {
      size_t nb = SL::img_bytes - o < 32768 ? SL::img_bytes - o : 32768;
      bulk_g2s(s_img + o, gi + o, (uint32_t)nb, bar_w);
    }
    const uint8_t* ga = p.blob + p.off_aux;
    for (size_t o = 0; o < aux_bytes; o += 32768) {
      size_t nb = aux_bytes - o < 32768 ? aux_bytes - o : 32768;
      bulk_g2s(reinterpret_cast<uint8_t*>(s_aux) + o, ga + o, (uint32_t)nb, bar_w);
    }
  }
  mbar_wait(bar_w, 0);

  const int n_groups = gridDim.x * 2;
  auto group_points = [&](int g) {
    int G = blockIdx.x * 2 + g;
    return G < p.n ? (p.n - 1 - G) / n_groups + 1 : 0;
  };

  if (warp == 0) {
    // =============================================================== MMA issuer
    if (!SELF && elect_one()) {
      const uint32_t idesc = make_idesc(BF16, TC_WIDTH);
      const uint32_t lbo = TC_WIDTH * 16, sbo = 128;
      const uint32_t img0 = smem_u32(s_img), bdyn0 = smem_u32(s_bdyn);
      // image segments [W0e | W1 | W2 | W3h | W3e | bias1 | bias2]: byte offsets computed, NOT
      // looked up in local arrays -- nvcc 12.9 overlapped two dynamically indexed constant local
      // arrays on the stack here (seg_off[1] and bias_off[1] shared a slot: wrong B operands)
      auto seg_off = [](int layer) { return layer == 0 ? 0u : (uint32_t)(KE + 128 * (layer - 1)) * 256u; };
      auto bias_off = [](int layer) { return (uint32_t)(KE + 384 + KE + 16 * (layer - 1)) * 256u; };
      constexpr uint32_t seg_w3e = (uint32_t)(KE + 384) * 256u;
      // running state per group: next layer, running tile index (its parity selects the A_e buffer)
      bool live[2] = {true, true};
      uint32_t ph[2] = {0u, 0u}, tile[2] = {0u, 0u};
      int layer_of[2] = {0, 0};
      while (live[0] || live[1]) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (!live[g]) continue;
          mbar_wait(bar_a + g, ph[g]);
          ph[g] ^= 1u;
          if (s_cnt[g * 8 + 5]) { live[g] = false; continue; }      // the group's last arrival
          tc_fence_after();
          const int layer = layer_of[g];
          const uint32_t tb = tmem_base + g * GRP_COLS;
          const uint32_t d_t = tb + COL_D;
          const uint32_t ae_t = tb + ((tile[g] & 1u) ? COL_AE1 : COL_AE);
          // bias block first (accumulate = 0 starts the tile from 1 * b_hi + 1 * b_lo)
          const uint32_t bsm = (layer == 0 || layer == 3)
                                   ? bdyn0 + (uint32_t)(g * 2 + (layer == 3 ? 1 : 0)) * 4096u
                                   : img0 + bias_off(layer);
          tc_mma_ts(d_t, tb + COL_ONE, make_b_desc(bsm, lbo, sbo), idesc, 0u);
          if (layer == 0) {
#pragma unroll
            for (int k = 0; k < KE / 16; ++k)
              tc_mma_ts(d_t, ae_t + k * 8,
                        make_b_desc(img0 + k * 2 * lbo, lbo, sbo), idesc, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              tc_mma_ts(d_t, tb + COL_AH + k * 8,
                        make_b_desc(img0 + seg_off(layer) + k * 2 * lbo, lbo, sbo), idesc, 1u);
            if (layer == 3) {
#pragma unroll
              for (int k = 0; k < KE / 16; ++k)
                tc_mma_ts(d_t, ae_t + k * 8,
                          make_b_desc(img0 + seg_w3e + k * 2 * lbo, lbo, sbo), idesc, 1u);
            }
          }
          tc_commit(bar_d + g);
          if (layer == 3) { layer_of[g] = 0; ++tile[g]; }
          else layer_of[g] = layer + 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================== workers
    const int g = (warp - 4) >> 2;            // group 0 | 1
    const int wq = warp & 3;                  // TMEM lane quarter this warp may access
    const int t = wq * 32 + lane;             // row of the tile == TMEM lane
    const int tg = t;                         // thread index inside the group
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t tb = tmem_base + g * GRP_COLS + lane_addr;
    float* e_s = s_e + g * 64;
    const float* Wx0 = s_aux + AUX_WX0;
    const float* Wx3 = Wx0 + NR_PAD * 128;
    // this thread's (n = tg, k = 0 | 1) pair of the group's two dynamic bias blocks
    uint32_t* bd0 = reinterpret_cast<uint32_t*>(s_bdyn + (size_t)(g * 2 + 0) * 4096 + (size_t)tg * 16);
    uint32_t* bd3 = reinterpret_cast<uint32_t*>(s_bdyn + (size_t)(g * 2 + 1) * 4096 + (size_t)tg * 16);
    const int G = blockIdx.x * 2 + g;
    uint32_t phd = 0u;
    int tile0 = 0;                              // tiles this group has run before the current point
    // SELF: this group's own MMA issue (same instruction sequence as the issuer warp's)
    const uint32_t w_idesc = make_idesc(BF16, TC_WIDTH);
    const uint32_t w_img0 = smem_u32(s_img), w_bdyn0 = smem_u32(s_bdyn);
    auto self_issue = [&](int layer, int aebuf) {
      const uint32_t lbo = TC_WIDTH * 16, sbo = 128;
      const uint32_t tbase = tmem_base + g * GRP_COLS;
      const uint32_t d_t = tbase + COL_D;
      const uint32_t ae_t = tbase + (aebuf ? COL_AE1 : COL_AE);
      const uint32_t seg = layer == 0 ? 0u : (uint32_t)(KE + 128 * (layer - 1)) * 256u;
      const uint32_t bsm = (layer == 0 || layer == 3)
                               ? w_bdyn0 + (uint32_t)(g * 2 + (layer == 3 ? 1 : 0)) * 4096u
                               : w_img0 + (uint32_t)(KE + 384 + KE + 16 * (layer - 1)) * 256u;
      tc_mma_ts(d_t, tbase + COL_ONE, make_b_desc(bsm, lbo, sbo), w_idesc, 0u);
      if (layer == 0) {
#pragma unroll
        for (int kk = 0; kk < KE / 16; ++kk)
          tc_mma_ts(d_t, ae_t + kk * 8, make_b_desc(w_img0 + kk * 2 * lbo, lbo, sbo), w_idesc, 1u);
      } else {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          tc_mma_ts(d_t, tbase + COL_AH + kk * 8, make_b_desc(w_img0 + seg + kk * 2 * lbo, lbo, sbo), w_idesc, 1u);
        if (layer == 3) {
#pragma unroll
          for (int kk = 0; kk < KE / 16; ++kk)
            tc_mma_ts(d_t, ae_t + kk * 8,
                      make_b_desc(w_img0 + (uint32_t)(KE + 384) * 256u + kk * 2 * lbo, lbo, sbo), w_idesc, 1u);
        }
      }
      tc_commit(bar_d + g);
    };
    // hand the A operand of `layer` over to the tensor core
    auto hand_over = [&](int layer, int aebuf) {
      tc_fence_before();
      if (SELF) {
        group_bar(3 + g);
        if (tg < 32 && elect_one()) { tc_fence_after(); self_issue(layer, aebuf); }
      } else {
        mbar_arrive(bar_a + g);
      }
    };
    {   // constant A operand of the bias block: columns (1, 1, 0, ..., 0), written once
      uint32_t one[8];
      one[0] = pack2<BF16, 0>(1.f, 1.f);
#pragma unroll
      for (int i = 1; i < 8; ++i) one[i] = 0u;
      TC_ST8(tb + COL_ONE, one);
      tc_wait_st();
    }

    for (int pt = G; pt < p.n; pt += n_groups) {
      // ---------------------------------------------- per-point (once per L lights)
      const f3 x = ld3(p.xyz + (size_t)pt * 3);
      const f3 xd = (KIND == NF_MLP_LVIS && p.xyz_dir) ? ld3(p.xyz_dir + (size_t)pt * 3) : x;
      f3 fr_t, fr_b, fr_n, v_loc;
      if (KIND == NF_MLP_LVIS) {
        if (tg < 3) e_s[tg] = (tg == 0 ? x.x : (tg == 1 ? x.y : x.z)) * p.xyz_scale;
        else if (tg < 3 + 3 * p.n_freqs_a) {
          int idx = tg - 3, f = idx / 3, c = idx % 3;
          float xv = (c == 0 ? x.x : (c == 1 ? x.y : x.z)) * p.xyz_scale;
          float s, co;
          sincosf(xv * (float)(1 << f), &s, &co);
          e_s[3 + 6 * f + c] = s;
          e_s[3 + 6 * f + 3 + c] = co;
        }
      } else {
        if (tg < p.z_dim) e_s[tg] = p.zlat[(size_t)pt * p.z_dim + tg];
        world2local_dev(ld3(p.normal + (size_t)pt * 3), fr_t, fr_b, fr_n);   // geom.py:119-149
        f3 v = l2n(ld3(p.cam + (size_t)pt * 3) - x, 1e-6f);                   // shape.py:137-144
        v_loc = mk3(dot3(fr_t, v), dot3(fr_b, v), dot3(fr_n, v));             // nerfactor.py:418
      }
      // fused rendering equation: this point's shading frame (every thread, same values)
      f3 n1 = mk3(0.f, 0.f, 1.f), v2 = n1, lam = n1;
      float a2 = 0.f, g_view = 0.f, cos_v = 0.f, abs_vn = 0.f;
      if (fuse) {
        n1 = l2n(l2n(ld3(p.f_normal + (size_t)pt * 3), 1e-6f), 1e-6f);     // nerfactor.py:212, microfacet.py:48
        v2 = l2n(l2n(ld3(p.f_cam + (size_t)pt * 3) - x, 1e-6f), 1e-6f);    // shape.py:137-144, microfacet.py:47
        const f3 alb = ld3(p.f_albedo + (size_t)pt * 3);
        lam = mk3(alb.x / NF_PI_F, alb.y / NF_PI_F, alb.z / NF_PI_F);
        const float rough = p.f_rough[pt];
        const float alpha = rough * rough;                                   // microfacet.py:54
        a2 = alpha * alpha;
        cos_v = dot3(n1, v2);                                                // :77
        const float cv2 = fminf(fmaxf(cos_v * cos_v, 0.f), 1.f);             // :82-84
        const float tan2 = fmaxf(divide_no_nan(1.f - cv2, cv2), 0.f);        // :85-87
        g_view = 2.f / (1.f + sqrtf(1.f + a2 * tan2));                       // :88-89
        abs_vn = fabsf(cos_v);
      }
      group_bar(1 + g);
      {
        // fold the per-point input columns into the biases of layer 0 and of the skip layer (fp32),
        // then hand them to the tensor core as an fp16 / bf16 hi + lo pair
        float a0 = s_aux[AUX_B + 0 * 128 + tg], a3 = s_aux[AUX_B + 3 * 128 + tg];
        for (int k = 0; k < p.nr; ++k) {
          float ev = e_s[k];
          a0 = fmaf(ev, Wx0[k * 128 + tg], a0);
          a3 = fmaf(ev, Wx3[k * 128 + tg], a3);
        }
        const uint32_t h0 = pack2<BF16, 0>(a0, 0.f), h3 = pack2<BF16, 0>(a3, 0.f);
        const float l0 = a0 - unpack_lo<BF16>(h0), l3 = a3 - unpack_lo<BF16>(h3);
        *bd0 = (h0 & 0xFFFFu) | (pack2<BF16, 0>(l0, 0.f) << 16);
        *bd3 = (h3 & 0xFFFFu) | (pack2<BF16, 0>(l3, 0.f) << 16);
        fence_proxy_async();
      }
      // tiles of this point: all lights (visibility) / only the front-lit ones (BRDF)
      int n_rows = p.L;
      int c0 = 0, c1 = 0, c2 = 0;                  // BRDF: cumulative front-lit counts of warps 0..2
      if (CULL) {
        // each warp scans a quarter of the lights, keeps the front-lit ones (l_loc.z > 0,
        // nerfactor.py:429-432) in increasing order and zeroes the others' output (:456-458);
        // visibility network: cos(shading normal, light) > -1e-5, a superset of the renderer's
        // cos > 0 (nerfactor.py:325-330) whatever the rounding of its own cosine
        f3 cn = fr_n;
        if (KIND == NF_MLP_LVIS) cn = l2n(l2n(ld3(p.cull_normal + (size_t)pt * 3), 1e-6f), 1e-6f);
        const int lq = (p.L + 3) / 4, l0 = wq * lq, l1 = min(p.L, l0 + lq);
        uint16_t* mylist = s_list + ((size_t)g * 4 + wq) * 256;
        int cnt = 0;
        for (int base = l0; base < l1; base += 32) {
          const int l = base + lane;
          bool lit = false;
          if (l < l1) {
            const float4 lp = s_lx[l];
            const f3 d = l2n(mk3(lp.x, lp.y, lp.z) - x, 1e-6f);
            lit = KIND == NF_MLP_LVIS ? dot3(cn, d) > -1e-5f : dot3(cn, d) > 0.f;
            if (!lit) p.out[(size_t)pt * p.L + l] = 0.f;
          }
          const unsigned m = __ballot_sync(0xffffffffu, lit);
          if (lit) mylist[cnt + __popc(m & ((1u << lane) - 1u))] = (uint16_t)l;
          cnt += __popc(m);
        }
        if (lane == 0) s_cnt[g * 8 + wq] = cnt;
        group_bar(1 + g);
        c0 = s_cnt[g * 8 + 0]; c1 = c0 + s_cnt[g * 8 + 1]; c2 = c1 + s_cnt[g * 8 + 2];
        n_rows = c2 + s_cnt[g * 8 + 3];
      }
      // BRDF network: at least one (possibly empty) tile, as before; visibility: none if unlit
      const int n_tiles = (n_rows > 0 || KIND == NF_MLP_BRDF) ? (n_rows > 0 ? (n_rows + 127) / 128 : 1) : 0;
      if (tg == 0) s_cnt[g * 8 + 4] = n_tiles;
      group_bar(1 + g);

      // Row of tile c handled by this thread, and its embedding -> A_e buffer (c & 1).
      struct RowInfo { int li; bool ok; float mask; f3 ldir; };
      auto embed_tile = [&](int c) {
        RowInfo r;
        r.li = c * 128 + t;                         // row of the point's (compacted) light list
        r.ok = r.li < n_rows;
        if (CULL) {
          int i = r.ok ? r.li : (n_rows > 0 ? n_rows - 1 : 0);
          const int seg = (i >= c0) + (i >= c1) + (i >= c2);
          i -= seg == 0 ? 0 : (seg == 1 ? c0 : (seg == 2 ? c1 : c2));
          r.li = n_rows > 0 ? (int)s_list[((size_t)g * 4 + seg) * 256 + i] : 0;
        }
        const int lc = CULL ? r.li : (r.li < p.L ? r.li : p.L - 1);
        r.mask = 1.f;
        float4 lp = s_lx[lc];
        f3 d = l2n(mk3(lp.x, lp.y, lp.z) - xd, 1e-6f);                        // shape.py:128-135
        r.ldir = d;
        float v[KE];
#pragma unroll
        for (int i = 0; i < KE; ++i) v[i] = 0.f;
        f3 q;
        int nf;
        if (KIND == NF_MLP_LVIS) { q = d; nf = 4; }
        else {
          f3 l_loc = mk3(dot3(fr_t, d), dot3(fr_b, d), dot3(fr_n, d));        // nerfactor.py:419
          r.mask = l_loc.z > 0.f ? 1.f : 0.f;                                 // :429-432
          q = dir2rusink_dev(l_loc, v_loc);                                   // geom.py:152-192
          nf = 2;
        }
        v[0] = q.x; v[1] = q.y; v[2] = q.z;
        float sx, cx, sy, cy, sz, cz;
        sincosf(q.x, &sx, &cx); sincosf(q.y, &sy, &cy); sincosf(q.z, &sz, &cz);
#pragma unroll
        for (int f = 0; f < (KIND == NF_MLP_LVIS ? 4 : 2); ++f) {
          if (f < nf) {
            v[3 + 6 * f + 0] = sx; v[3 + 6 * f + 1] = sy; v[3 + 6 * f + 2] = sz;
            v[3 + 6 * f + 3] = cx; v[3 + 6 * f + 4] = cy; v[3 + 6 * f + 5] = cz;
            float nsx = 2.f * sx * cx, ncx = 1.f - 2.f * sx * sx;
            float nsy = 2.f * sy * cy, ncy = 1.f - 2.f * sy * sy;
            float nsz = 2.f * sz * cz, ncz = 1.f - 2.f * sz * sz;
            sx = nsx; cx = ncx; sy = nsy; cy = ncy; sz = nsz; cz = ncz;
          }
        }
        uint32_t pk[KE / 2];
#pragma unroll
        for (int i = 0; i < KE / 2; ++i) pk[i] = pack2<BF16, 0>(v[2 * i], v[2 * i + 1]);
        const uint32_t ae = tb + (((tile0 + c) & 1) ? COL_AE1 : COL_AE);
        if (KE == 32) { TC_ST16(ae, pk); }
        else { TC_ST8(ae, pk); }
        tc_wait_st();
        return r;
      };

      RowInfo cur;
      if (n_tiles > 0) {
        cur = embed_tile(0);
        hand_over(0, tile0 & 1);
      }
      for (int c = 0; c < n_tiles; ++c) {
        // ------------------------------------------------ layers 0..2: ReLU + 16-bit -> A_h
        for (int layer = 0; layer < 3; ++layer) {
          mbar_wait(bar_d + g, phd);
          phd ^= 1u;
          tc_fence_after();
          uint32_t ra[32], rb[32], rc[32], rd[32];
          TC_LD32(ra, tb + COL_D);
          TC_LD32(rb, tb + COL_D + 32);
          tc_wait_ld();
          TC_LD32(rc, tb + COL_D + 64);       // second half in flight while the first converts
          TC_LD32(rd, tb + COL_D + 96);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack2<BF16, 1>(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]));
          TC_ST16(tb + COL_AH, pk);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack2<BF16, 1>(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]));
          TC_ST16(tb + COL_AH + 16, pk);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack2<BF16, 1>(__uint_as_float(rc[2 * i]), __uint_as_float(rc[2 * i + 1]));
          TC_ST16(tb + COL_AH + 32, pk);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack2<BF16, 1>(__uint_as_float(rd[2 * i]), __uint_as_float(rd[2 * i + 1]));
          TC_ST16(tb + COL_AH + 48, pk);
          tc_wait_st();
          hand_over(layer + 1, (tile0 + c) & 1);
        }
        // ------------------------------------------------ layer 3 is in flight: the NEXT tile's
        // per-row embedding goes into the other A_e buffer now (off the critical chain)
        // (visibility network only: with the group-issued MMAs of the compacted BRDF network the
        // reordering measured slower, 28.5 vs 27.2 ms)
        constexpr bool PIPE = KIND == NF_MLP_LVIS;
        const bool has_next = PIPE && c + 1 < n_tiles;
        RowInfo nxt = cur;
        if (has_next) nxt = embed_tile(c + 1);
        // ------------------------------------------------ layer 3 done: accumulator -> registers,
        // hand the next tile's layer 0 over, THEN the head (it only needs the registers)
        mbar_wait(bar_d + g, phd);
        phd ^= 1u;
        tc_fence_after();
        float acc0 = 0.f, acc1 = 0.f;
        {
          uint32_t ra[32], rb[32], rc[32], rd[32];
          TC_LD32(ra, tb + COL_D);
          TC_LD32(rb, tb + COL_D + 32);
          TC_LD32(rc, tb + COL_D + 64);
          TC_LD32(rd, tb + COL_D + 96);
          tc_wait_ld();
          if (has_next) hand_over(0, (tile0 + c + 1) & 1);
          const float4* wo = reinterpret_cast<const float4*>(s_aux + AUX_WOUT);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w0 = wo[i], w1 = wo[8 + i];
            acc0 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 0]), 0.f), w0.x, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 1]), 0.f), w0.y, acc1);
            acc0 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 2]), 0.f), w0.z, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 3]), 0.f), w0.w, acc1);
            acc0 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 0]), 0.f), w1.x, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 1]), 0.f), w1.y, acc1);
            acc0 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 2]), 0.f), w1.z, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 3]), 0.f), w1.w, acc1);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w0 = wo[16 + i], w1 = wo[24 + i];
            acc0 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 0]), 0.f), w0.x, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 1]), 0.f), w0.y, acc1);
            acc0 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 2]), 0.f), w0.z, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 3]), 0.f), w0.w, acc1);
            acc0 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 0]), 0.f), w1.x, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 1]), 0.f), w1.y, acc1);
            acc0 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 2]), 0.f), w1.z, acc0);
            acc1 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 3]), 0.f), w1.w, acc1);
          }
        }
        const int li = cur.li;
        const bool row_ok = cur.ok;
        const f3 ldir = cur.ldir;
        float o = (acc0 + acc1) + s_aux[AUX_BOUT];
        o = apply_act(p.out_act, o) * cur.mask;
        if (row_ok && p.out) p.out[(size_t)pt * p.L + li] = o;
        if (fuse) {
          // one term of the rendering equation (nerfactor.py:325-336) with the GGX lobe of
          // brdf/microfacet/microfacet.py:30-111 (same reduction as nf_integrate.cu eval_pair)
          float q0 = 0.f, q1 = 0.f, q2 = 0.f;
          if (row_ok) {
            const float cosl = dot3(ldir, n1);                               // :325
            const float wgt = (cosl > 0.f ? o : 0.f) * cosl;                 // :329-335 (area in s_lrgb)
            const f3 hs = ldir + v2;
            const f3 h = hs * rsqrtf(fmaxf(dot3(hs, hs), 1e-6f));            // microfacet.py:51-52
            const float om = 1.f - dot3(ldir, h);
            const float om2 = om * om;
            const float fr = p.f_f0 + (1.f - p.f_f0) * (om2 * om2 * om);     // :106-111
            const float cm = dot3(h, n1);                                    // :96
            const float cm2 = cm * cm;
            const float u = fmaf(a2, cm2, 1.f - cm2);
            const bool on = (cm > 0.f) && (dot3(h, v2) * cos_v > 0.f);       // chi_d, chi_g
            const float den = (4.f * NF_PI_F) * (u * u) * (fabsf(cosl) * abs_vn);
            const float spec = (on && den != 0.f) ? __fdividef(fr * (g_view * a2), den) : 0.f;
            const float4 lc = s_lrgb[li];
            q0 = (spec + lam.x) * (wgt * lc.x);
            q1 = (spec + lam.y) * (wgt * lc.y);
            q2 = (spec + lam.z) * (wgt * lc.z);
          }
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) {
            q0 += __shfl_xor_sync(0xffffffffu, q0, sft);
            q1 += __shfl_xor_sync(0xffffffffu, q1, sft);
            q2 += __shfl_xor_sync(0xffffffffu, q2, sft);
          }
          if (lane == 0) {          // fixed slot per (chunk, warp): summed in a fixed order below
            float* slot = s_red + ((size_t)g * 16 + c * 4 + wq) * 4;
            slot[0] = q0; slot[1] = q1; slot[2] = q2;
          }
        }
        if (!PIPE && c + 1 < n_tiles) {          // plain order: embedding, then layer 0
          nxt = embed_tile(c + 1);
          hand_over(0, (tile0 + c + 1) & 1);
        }
        cur = nxt;
      }
      tile0 += n_tiles;
      if (fuse) {
        group_bar(1 + g);
        if (tg == 0) {
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
          for (int q = 0; q < n_tiles * 4; ++q) {
            const float* slot = s_red + ((size_t)g * 16 + q) * 4;
            r0 += slot[0]; r1 += slot[1]; r2 += slot[2];
          }
          float* orgb = p.f_rgb + (size_t)pt * 3;
          r0 = fminf(fmaxf(r0, 0.f), 1.f); r1 = fminf(fmaxf(r1, 0.f), 1.f); r2 = fminf(fmaxf(r2, 0.f), 1.f);
          orgb[0] = p.f_srgb ? linear2srgb_dev(r0) : r0;                     // nerfactor.py:338-339
          orgb[1] = p.f_srgb ? linear2srgb_dev(r1) : r1;
          orgb[2] = p.f_srgb ? linear2srgb_dev(r2) : r2;
        }
      }
    }
    if (!SELF) {                    // tell the issuer warp that this group has no more tiles
      if (tg == 0) s_cnt[g * 8 + 5] = 1;
      mbar_arrive(bar_a + g);
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================
// Version 3 of the light-visibility kernel: everything that happens once per surface POINT is
// taken off the workers' critical path, and a tile is 128 rows of the group's ROW STREAM, not of
// one point.
// Measured on B200 (profiles/r2_k2_analysis.md, section 4): in version 2 the per-point phase --
// 63-term fold of the positional encoding of xyz into two 128-wide biases, hi/lo split, the first
// tile's per-row embedding -- costs about as much as 1.4 tiles and both worker groups sit in it
// with the tensor pipe idle (cutting the tiles per point from 4 to 3 by front-lit culling did not
// shorten that kernel at all).  Here
//   * warps 1 and 2 (idle so far) are PREFETCHERS, one per worker group: they run ahead of the
//     group, fold the next points' biases, build the front-lit light list (CULL) and publish a
//     small record per point; four record buffers per group with full / free mbarriers;
//   * the rows of a group form one stream -- the (front-lit) lights of its points one after the
//     other -- cut into tiles of 128: a tile may end one point and start the next (at most two
//     points per tile), so ~260 front-lit lights cost 2.03 tiles per point instead of 3;
//   * the per-point biases of layer 0 and of the skip layer sit in ONE 2 KB block per layer: point
//     number i owns rows k = 2 (i % 4), 2 (i % 4) + 1 (hi, lo) of it, and every tile row selects
//     its point through its own "ones" operand ((1, 1) at that position, zeros elsewhere), written
//     together with the row's light-direction embedding.  The static bias blocks of layers 1, 2
//     repeat (hi, lo) at all four positions.  k = 8..15 is a shared 2 KB block of zeros reached
//     through the descriptor's leading-dimension offset;
//   * the embedding of the next tile is written under layer 3 of the current one, also across a
//     point boundary; the issuer serves hand-overs and needs no per-tile information.
// Light positions are read from global memory (L1-resident 12 KB) instead of a shared table.
template <int KIND>
struct SmemLayout3 {
  static constexpr int KE = KindCfg<KIND>::KE;
  static constexpr int NR_PAD = KindCfg<KIND>::NR_PAD;
  static constexpr int LMAX = 1024;
  static constexpr int NBUF = 4;                                 // point records in flight per group
  static constexpr int REC_F = KIND == NF_MLP_BRDF ? 16 : 4;     // floats per point record
  static constexpr size_t img_bytes = ((size_t)img_halves<KIND>() + 2 * 16 * 128) * 2;
  static constexpr size_t aux_floats = AUX_WX0 + 2 * (size_t)NR_PAD * 128;
  static constexpr size_t off_img = 0;
  static constexpr size_t off_bdyn = img_bytes;                  // [2 g][2 layer] x 2048 B
  static constexpr size_t off_zero = off_bdyn + 4 * 2048;        // 2048 B of zeros (k = 8..15)
  static constexpr size_t off_aux = off_zero + 2048;
  static constexpr size_t off_rec = off_aux + aux_floats * 4;    // [2 g][NBUF] records (REC_F floats)
  static constexpr size_t off_list = off_rec + 2 * NBUF * REC_F * 4;   // [2 g][NBUF][LMAX] u16
  static constexpr size_t off_es = off_list + 2 * NBUF * (size_t)LMAX * 2;   // [2 prefetchers][64] f32
  static constexpr size_t off_bar = off_es + 2 * 64 * 4;         // 5 + 16 barriers + tmem pointer
  static constexpr size_t off_flag = off_bar + 192;              // int: done[2]
  static constexpr size_t total = off_flag + 16;
};
static_assert(SmemLayout3<NF_MLP_LVIS>::total <= 232448 && SmemLayout3<NF_MLP_BRDF>::total <= 232448,
              "shared memory budget");
constexpr int COL_ONE1 = 232;      // second "ones" operand (tiles alternate, like COL_AE / COL_AE1)

// KIND = NF_MLP_BRDF (always CULL): the per-point record carries the shading frame and the local view
// direction, the fold runs over the latent code z, the per-row embedding is the Rusinkiewicz triple
// (nerfactor.py:413-457) -- everything else is the same machinery.
template <int KIND, int BF16, int CULL>
__device__ __forceinline__ void tc3_body(const TcParams& p) {
  static_assert(KIND != NF_MLP_BRDF || CULL == 1, "the BRDF network always runs on the front-lit lights");
  using SL = SmemLayout3<KIND>;
  constexpr int KE = SL::KE;
  constexpr int NR_PAD = SL::NR_PAD;
  constexpr int NBUF = SL::NBUF;
  constexpr int REC_F = SL::REC_F;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_img = smem + SL::off_img;
  uint8_t* s_bdyn = smem + SL::off_bdyn;
  float* s_aux = reinterpret_cast<float*>(smem + SL::off_aux);
  float* s_rec = reinterpret_cast<float*>(smem + SL::off_rec);
  uint16_t* s_list = reinterpret_cast<uint16_t*>(smem + SL::off_list);
  float* s_es = reinterpret_cast<float*>(smem + SL::off_es);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL::off_bar);
  uint64_t* bar_w = bars + 0;          // weights landed
  uint64_t* bar_a = bars + 1;          // [2] A operand ready (128 arrivals)
  uint64_t* bar_d = bars + 3;          // [2] D accumulator ready (tcgen05.commit)
  uint64_t* rec_full = bars + 5;       // [2 g][NBUF] record + biases of a point published
  uint64_t* rec_free = bars + 5 + 2 * NBUF;   // [2 g][NBUF] the group is done with that point
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 5 + 4 * NBUF);
  volatile int* s_flag = reinterpret_cast<volatile int*>(smem + SL::off_flag);   // [g] done

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---------------------------------------------------------------- set-up
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a + 0, 128); mbar_init(bar_a + 1, 128);
    mbar_init(bar_d + 0, 1); mbar_init(bar_d + 1, 1);
    for (int i = 0; i < 2 * NBUF; ++i) { mbar_init(rec_full + i, 1); mbar_init(rec_free + i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (4 * 2048 + 2048) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(s_bdyn)[i] = 0u;
  if (threadIdx.x < 4) s_flag[threadIdx.x] = 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t aux_bytes = (uint32_t)(SL::aux_floats * 4);
    mbar_expect_tx(bar_w, (uint32_t)SL::img_bytes + aux_bytes);
    const uint8_t* gi = p.blob + p.off_img;
    for (size_t o = 0; o < SL::img_bytes; o += 32768) {
      size_t nb = SL::img_bytes - o < 32768 ? SL::img_bytes - o : 32768;
      bulk_g2s(s_img + o, gi + o, (uint32_t)nb, bar_w);
    }
    const uint8_t* ga = p.blob + p.off_aux;
    for (size_t o = 0; o < aux_bytes; o += 32768) {
      size_t nb = aux_bytes - o < 32768 ? aux_bytes - o : 32768;
      bulk_g2s(reinterpret_cast<uint8_t*>(s_aux) + o, ga + o, (uint32_t)nb, bar_w);
    }
  }
  mbar_wait(bar_w, 0);
  const int n_groups = gridDim.x * 2;

  if (warp == 0) {
    // =============================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(BF16, TC_WIDTH);
      const uint32_t lbo = TC_WIDTH * 16, sbo = 128;
      const uint32_t img0 = smem_u32(s_img), bdyn0 = smem_u32(s_bdyn);
      const uint32_t zero0 = smem_u32(smem + SL::off_zero);
      auto seg_off = [](int layer) { return layer == 0 ? 0u : (uint32_t)(KE + 128 * (layer - 1)) * 256u; };
      auto bias_off = [](int layer) { return (uint32_t)(KE + 384 + KE + 16 * (layer - 1)) * 256u; };
      constexpr uint32_t seg_w3e = (uint32_t)(KE + 384) * 256u;
      bool live[2] = {true, true};
      uint32_t ph[2] = {0u, 0u}, tile[2] = {0u, 0u};
      int layer_of[2] = {0, 0};
      while (live[0] || live[1]) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (!live[g]) continue;
          mbar_wait(bar_a + g, ph[g]);
          ph[g] ^= 1u;
          if (s_flag[g]) { live[g] = false; continue; }      // the group's last arrival
          tc_fence_after();
          const int layer = layer_of[g];
          const uint32_t tb = tmem_base + g * GRP_COLS;
          const uint32_t d_t = tb + COL_D;
          const uint32_t ae_t = tb + ((tile[g] & 1u) ? COL_AE1 : COL_AE);
          const uint32_t one_t = tb + ((tile[g] & 1u) ? COL_ONE1 : COL_ONE);
          // bias block first (accumulate = 0 starts the tile from 1 * b_hi + 1 * b_lo)
          if (layer == 0 || layer == 3) {
            const uint32_t bsm = bdyn0 + (uint32_t)(g * 2 + (layer == 3 ? 1 : 0)) * 2048u;
            tc_mma_ts(d_t, one_t, make_b_desc(bsm, zero0 - bsm, sbo), idesc, 0u);
          } else {
            tc_mma_ts(d_t, one_t, make_b_desc(img0 + bias_off(layer), lbo, sbo), idesc, 0u);
          }
          if (layer == 0) {
#pragma unroll
            for (int k = 0; k < KE / 16; ++k)
              tc_mma_ts(d_t, ae_t + k * 8, make_b_desc(img0 + k * 2 * lbo, lbo, sbo), idesc, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              tc_mma_ts(d_t, tb + COL_AH + k * 8,
                        make_b_desc(img0 + seg_off(layer) + k * 2 * lbo, lbo, sbo), idesc, 1u);
            if (layer == 3) {
#pragma unroll
              for (int k = 0; k < KE / 16; ++k)
                tc_mma_ts(d_t, ae_t + k * 8,
                          make_b_desc(img0 + seg_w3e + k * 2 * lbo, lbo, sbo), idesc, 1u);
            }
          }
          tc_commit(bar_d + g);
          if (layer == 3) { layer_of[g] = 0; ++tile[g]; }
          else layer_of[g] = layer + 1;
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ============================================================ per-point prefetcher of group g
    const int g = warp - 1;
    const int G = blockIdx.x * 2 + g;
    float* es = s_es + g * 64;
    const float* Wx0 = s_aux + AUX_WX0;
    const float* Wx3 = Wx0 + NR_PAD * 128;
    int i = 0;
    for (int pt = G; pt < p.n; pt += n_groups, ++i) {
      const int b = i & (NBUF - 1);
      if (i >= NBUF) mbar_wait(rec_free + g * NBUF + b, (uint32_t)((i / NBUF) - 1) & 1u);
      const f3 x = ld3(p.xyz + (size_t)pt * 3);
      f3 fr_t = mk3(1.f, 0.f, 0.f), fr_b = mk3(0.f, 1.f, 0.f), fr_n = mk3(0.f, 0.f, 1.f), v_loc = fr_n;
      if (KIND == NF_MLP_LVIS) {
        // positional encoding of xyz (embedder.py:46-47), fp32
        if (lane < 3) es[lane] = (lane == 0 ? x.x : (lane == 1 ? x.y : x.z)) * p.xyz_scale;
        for (int idx = lane; idx < 3 * p.n_freqs_a; idx += 32) {
          const int fq = idx / 3, c = idx % 3;
          const float xv = (c == 0 ? x.x : (c == 1 ? x.y : x.z)) * p.xyz_scale;
          float sn, cs;
          sincosf(xv * (float)(1 << fq), &sn, &cs);
          es[3 + 6 * fq + c] = sn;
          es[3 + 6 * fq + 3 + c] = cs;
        }
      } else {
        // latent code z of the point; shading frame and local view direction
        if (lane < p.z_dim) es[lane] = p.zlat[(size_t)pt * p.z_dim + lane];
        world2local_dev(ld3(p.normal + (size_t)pt * 3), fr_t, fr_b, fr_n);   // geom.py:119-149
        const f3 v = l2n(ld3(p.cam + (size_t)pt * 3) - x, 1e-6f);             // shape.py:137-144
        v_loc = mk3(dot3(fr_t, v), dot3(fr_b, v), dot3(fr_n, v));             // nerfactor.py:418
      }
      __syncwarp();
      // fold the per-point input columns into the biases of layer 0 and of the skip layer (fp32):
      // four output columns per lane
      float4 a0 = *reinterpret_cast<const float4*>(s_aux + AUX_B + 0 * 128 + 4 * lane);
      float4 a3 = *reinterpret_cast<const float4*>(s_aux + AUX_B + 3 * 128 + 4 * lane);
#pragma unroll 4
      for (int k = 0; k < p.nr; ++k) {
        const float ev = es[k];
        const float4 w0 = *reinterpret_cast<const float4*>(Wx0 + k * 128 + 4 * lane);
        const float4 w3 = *reinterpret_cast<const float4*>(Wx3 + k * 128 + 4 * lane);
        a0.x = fmaf(ev, w0.x, a0.x); a0.y = fmaf(ev, w0.y, a0.y);
        a0.z = fmaf(ev, w0.z, a0.z); a0.w = fmaf(ev, w0.w, a0.w);
        a3.x = fmaf(ev, w3.x, a3.x); a3.y = fmaf(ev, w3.y, a3.y);
        a3.z = fmaf(ev, w3.z, a3.z); a3.w = fmaf(ev, w3.w, a3.w);
      }
      __syncwarp();                    // es is rewritten for the next point
      // 16-bit hi + lo pair per column into rows k = 2 b, 2 b + 1 of the group's bias block of
      // that layer (element (n, k) at n 16 + k 2 bytes)
      uint8_t* blk0 = s_bdyn + (size_t)(g * 2 + 0) * 2048 + (size_t)b * 4;
      uint8_t* blk3 = s_bdyn + (size_t)(g * 2 + 1) * 2048 + (size_t)b * 4;
      const float v0[4] = {a0.x, a0.y, a0.z, a0.w}, v3[4] = {a3.x, a3.y, a3.z, a3.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = 4 * lane + j;
        const uint32_t h0 = pack2<BF16, 0>(v0[j], 0.f), h3 = pack2<BF16, 0>(v3[j], 0.f);
        const float l0 = v0[j] - unpack_lo<BF16>(h0), l3 = v3[j] - unpack_lo<BF16>(h3);
        *reinterpret_cast<uint32_t*>(blk0 + (size_t)col * 16) = (h0 & 0xFFFFu) | (pack2<BF16, 0>(l0, 0.f) << 16);
        *reinterpret_cast<uint32_t*>(blk3 + (size_t)col * 16) = (h3 & 0xFFFFu) | (pack2<BF16, 0>(l3, 0.f) << 16);
      }
      int n_rows = p.L;
      if (CULL) {
        // front-lit lights in increasing order; the others' output is zero (nerfactor.py:329-330).
        // cos(shading normal, light) > -1e-5: a superset of the renderer's cos > 0 whatever the
        // rounding of its own cosine
        // (BRDF network: l_loc.z > 0 exactly, nerfactor.py:429-432, unlit outputs zero :456-458)
        const f3 cn = KIND == NF_MLP_BRDF ? fr_n
                                          : l2n(l2n(ld3(p.cull_normal + (size_t)pt * 3), 1e-6f), 1e-6f);
        uint16_t* list = s_list + (size_t)(g * NBUF + b) * SL::LMAX;
        int cnt = 0;
        for (int base = 0; base < p.L; base += 32) {
          const int l = base + lane;
          bool lit = false;
          if (l < p.L) {
            const f3 d = l2n(ld3(p.lxyz + (size_t)l * 3) - x, 1e-6f);
            lit = KIND == NF_MLP_BRDF ? dot3(cn, d) > 0.f : dot3(cn, d) > -1e-5f;
            if (!lit) p.out[(size_t)pt * p.L + l] = 0.f;
          }
          const unsigned m = __ballot_sync(0xffffffffu, lit);
          if (lit) list[cnt + __popc(m & ((1u << lane) - 1u))] = (uint16_t)l;
          cnt += __popc(m);
        }
        n_rows = cnt;
      }
      if (lane == 0) {
        const f3 xd = (KIND == NF_MLP_LVIS && p.xyz_dir) ? ld3(p.xyz_dir + (size_t)pt * 3) : x;
        float* rec = s_rec + (size_t)(g * NBUF + b) * REC_F;
        rec[0] = xd.x; rec[1] = xd.y; rec[2] = xd.z; rec[3] = __int_as_float(n_rows);
        if (KIND == NF_MLP_BRDF) {
          rec[4] = fr_t.x; rec[5] = fr_t.y; rec[6] = fr_t.z;
          rec[7] = fr_b.x; rec[8] = fr_b.y; rec[9] = fr_b.z;
          rec[10] = fr_n.x; rec[11] = fr_n.y; rec[12] = fr_n.z;
          rec[13] = v_loc.x; rec[14] = v_loc.y; rec[15] = v_loc.z;
        }
      }
      fence_proxy_async();           // the bias blocks are read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(rec_full + g * NBUF + b);
    }
  } else if (warp >= 4) {
    // ================================================================== workers
    const int g = (warp - 4) >> 2;            // group 0 | 1
    const int wq = warp & 3;                  // TMEM lane quarter this warp may access
    const int t = wq * 32 + lane;             // row of the tile == TMEM lane == thread of the group
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t tb = tmem_base + g * GRP_COLS + lane_addr;
    const int G = blockIdx.x * 2 + g;
    uint32_t phd = 0u;
    struct PointState { int pt, b, n_rows; f3 xd, fr_t, fr_b, fr_n, v_loc; };
    // a tile: rows [a_start, a_start + a_cnt) of point a, then the first b_cnt rows of the next point
    struct Tile { PointState a, b; int a_start, a_cnt, b_cnt; bool a_last, b_last; };
    struct RowInfo { int pt, li; bool ok; };
    // record of the group's i-th point (published by the prefetcher); false past the last point.
    // Waiting again for a record that is still held returns at once.
    auto acquire = [&](int i, PointState& s) {
      s.pt = G + i * n_groups;
      if (s.pt >= p.n) return false;
      s.b = i & (NBUF - 1);
      mbar_wait(rec_full + g * NBUF + s.b, (uint32_t)(i / NBUF) & 1u);
      const float* rec = s_rec + (size_t)(g * NBUF + s.b) * REC_F;
      s.xd = mk3(rec[0], rec[1], rec[2]);
      s.n_rows = __float_as_int(rec[3]);
      if (KIND == NF_MLP_BRDF) {
        s.fr_t = mk3(rec[4], rec[5], rec[6]);
        s.fr_b = mk3(rec[7], rec[8], rec[9]);
        s.fr_n = mk3(rec[10], rec[11], rec[12]);
        s.v_loc = mk3(rec[13], rec[14], rec[15]);
      }
      return true;
    };
    // head of the row stream: point number hp, of which hoff rows are consumed (hs valid if hvalid)
    int hp = 0, hoff = 0;
    bool hvalid = false;
    PointState hs;
    // next tile of the stream: 1 = built, 0 = the stream has ended, -1 = `overlap` was set and the
    // next point has no row (it can only be released once every thread has seen that, which needs
    // a barrier the caller does not want under a running layer)
    auto build = [&](bool overlap, Tile& T) {
      for (;;) {
        if (!hvalid) {
          if (!acquire(hp, hs)) return 0;
          hvalid = true;
          hoff = 0;
        }
        if (hs.n_rows > hoff) break;
        if (overlap) return -1;
        group_bar(1 + g);                               // every thread has read the record
        if (t == 0) mbar_arrive(rec_free + g * NBUF + hs.b);
        ++hp;
        hvalid = false;
      }
      T.a = hs; T.b = hs;
      T.a_start = hoff;
      T.a_cnt = min(128, hs.n_rows - hoff);
      T.b_cnt = 0;
      T.a_last = false; T.b_last = false;
      hoff += T.a_cnt;
      if (hoff == hs.n_rows) {                          // this tile ends point a
        T.a_last = true;
        ++hp;
        hvalid = false;
        if (T.a_cnt < 128 && acquire(hp, hs)) {         // room for the first rows of the next one
          hvalid = true;
          hoff = 0;
          if (hs.n_rows > 0) {
            T.b = hs;
            T.b_cnt = min(128 - T.a_cnt, hs.n_rows);
            hoff = T.b_cnt;
            if (hoff == hs.n_rows) { T.b_last = true; ++hp; hvalid = false; }
          }     // (a point without rows stays the head: the next build releases it)
        }
      }
      return 1;
    };
    // row `t` of tile T: its light-direction embedding and its "ones" operand -> buffers `par`
    auto embed_tile = [&](const Tile& T, int par) {
      RowInfo r;
      const bool in_a = t < T.a_cnt;
      const PointState& s = in_a ? T.a : T.b;
      r.ok = t < T.a_cnt + T.b_cnt;
      int row = in_a ? T.a_start + t : t - T.a_cnt;
      if (!r.ok) row = T.b_cnt > 0 ? T.b_cnt - 1 : T.a_start + T.a_cnt - 1;     // any valid row
      r.pt = s.pt;
      r.li = CULL ? (int)s_list[(size_t)(g * NBUF + s.b) * SL::LMAX + row] : row;
      const f3 d = l2n(ld3(p.lxyz + (size_t)r.li * 3) - s.xd, 1e-6f);            // shape.py:128-135
      f3 q = d;
      if (KIND == NF_MLP_BRDF) {
        const f3 l_loc = mk3(dot3(s.fr_t, d), dot3(s.fr_b, d), dot3(s.fr_n, d));  // nerfactor.py:419
        q = dir2rusink_dev(l_loc, s.v_loc);                                       // geom.py:152-192
      }
      float v[KE];
#pragma unroll
      for (int i = 0; i < KE; ++i) v[i] = 0.f;
      v[0] = q.x; v[1] = q.y; v[2] = q.z;
      float sx, cx, sy, cy, sz, cz;
      sincosf(q.x, &sx, &cx); sincosf(q.y, &sy, &cy); sincosf(q.z, &sz, &cz);
#pragma unroll
      for (int fq = 0; fq < (KIND == NF_MLP_LVIS ? 4 : 2); ++fq) {
        v[3 + 6 * fq + 0] = sx; v[3 + 6 * fq + 1] = sy; v[3 + 6 * fq + 2] = sz;
        v[3 + 6 * fq + 3] = cx; v[3 + 6 * fq + 4] = cy; v[3 + 6 * fq + 5] = cz;
        const float nsx = 2.f * sx * cx, ncx = 1.f - 2.f * sx * sx;
        const float nsy = 2.f * sy * cy, ncy = 1.f - 2.f * sy * sy;
        const float nsz = 2.f * sz * cz, ncz = 1.f - 2.f * sz * sz;
        sx = nsx; cx = ncx; sy = nsy; cy = ncy; sz = nsz; cz = ncz;
      }
      uint32_t pk[KE / 2];
#pragma unroll
      for (int i = 0; i < KE / 2; ++i) pk[i] = pack2<BF16, 0>(v[2 * i], v[2 * i + 1]);
      if (KE == 32) { TC_ST16(tb + (par ? COL_AE1 : COL_AE), pk); }
      else { TC_ST8(tb + (par ? COL_AE1 : COL_AE), pk); }
      // (1, 1) at k = 2 b, 2 b + 1: this row takes the bias pair of ITS point
      uint32_t one[8];
      const uint32_t pair = pack2<BF16, 0>(1.f, 1.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) one[i] = (i == s.b) ? pair : 0u;
      TC_ST8(tb + (par ? COL_ONE1 : COL_ONE), one);
      tc_wait_st();
      return r;
    };
    auto hand_over = [&]() {
      tc_fence_before();
      mbar_arrive(bar_a + g);
    };

    int tile = 0;
    Tile cur;
    RowInfo rcur;
    bool have = build(false, cur) == 1;
    if (have) {
      rcur = embed_tile(cur, 0);
      hand_over();
    }
    while (have) {
      // ------------------------------------------------ layers 0..2: ReLU + 16-bit -> A_h
      for (int layer = 0; layer < 3; ++layer) {
        mbar_wait(bar_d + g, phd);
        phd ^= 1u;
        tc_fence_after();
        uint32_t ra[32], rb[32], rc[32], rd[32];
        TC_LD32(ra, tb + COL_D);
        TC_LD32(rb, tb + COL_D + 32);
        tc_wait_ld();
        TC_LD32(rc, tb + COL_D + 64);       // second half in flight while the first converts
        TC_LD32(rd, tb + COL_D + 96);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack2<BF16, 1>(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]));
        TC_ST16(tb + COL_AH, pk);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack2<BF16, 1>(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]));
        TC_ST16(tb + COL_AH + 16, pk);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack2<BF16, 1>(__uint_as_float(rc[2 * i]), __uint_as_float(rc[2 * i + 1]));
        TC_ST16(tb + COL_AH + 32, pk);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack2<BF16, 1>(__uint_as_float(rd[2 * i]), __uint_as_float(rd[2 * i + 1]));
        TC_ST16(tb + COL_AH + 48, pk);
        tc_wait_st();
        hand_over();
      }
      // ------------------------------------------------ layer 3 is in flight: the NEXT tile of the
      // stream gets its embedding now
      Tile nx = cur;
      int st = build(true, nx);
      RowInfo rn = rcur;
      if (st == 1) rn = embed_tile(nx, (tile + 1) & 1);
      // ------------------------------------------------ layer 3 done: accumulator -> registers,
      // hand the next tile's layer 0 over, THEN the head (it only needs the registers)
      mbar_wait(bar_d + g, phd);
      phd ^= 1u;
      tc_fence_after();
      float acc0 = 0.f, acc1 = 0.f;
      {
        uint32_t ra[32], rb[32], rc[32], rd[32];
        TC_LD32(ra, tb + COL_D);
        TC_LD32(rb, tb + COL_D + 32);
        TC_LD32(rc, tb + COL_D + 64);
        TC_LD32(rd, tb + COL_D + 96);
        tc_wait_ld();
        // every MMA that reads the bias rows of a point this tile ends has completed
        if (t == 0) {
          if (cur.a_last) mbar_arrive(rec_free + g * NBUF + cur.a.b);
          if (cur.b_cnt > 0 && cur.b_last) mbar_arrive(rec_free + g * NBUF + cur.b.b);
        }
        if (st == 1) hand_over();
        const float4* wo = reinterpret_cast<const float4*>(s_aux + AUX_WOUT);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w0 = wo[i], w1 = wo[8 + i];
          acc0 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 0]), 0.f), w0.x, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 1]), 0.f), w0.y, acc1);
          acc0 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 2]), 0.f), w0.z, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(ra[4 * i + 3]), 0.f), w0.w, acc1);
          acc0 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 0]), 0.f), w1.x, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 1]), 0.f), w1.y, acc1);
          acc0 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 2]), 0.f), w1.z, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(rb[4 * i + 3]), 0.f), w1.w, acc1);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w0 = wo[16 + i], w1 = wo[24 + i];
          acc0 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 0]), 0.f), w0.x, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 1]), 0.f), w0.y, acc1);
          acc0 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 2]), 0.f), w0.z, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(rc[4 * i + 3]), 0.f), w0.w, acc1);
          acc0 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 0]), 0.f), w1.x, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 1]), 0.f), w1.y, acc1);
          acc0 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 2]), 0.f), w1.z, acc0);
          acc1 = fmaf(fmaxf(__uint_as_float(rd[4 * i + 3]), 0.f), w1.w, acc1);
        }
      }
      const float o = apply_act(p.out_act, (acc0 + acc1) + s_aux[AUX_BOUT]);
      if (rcur.ok) p.out[(size_t)rcur.pt * p.L + rcur.li] = o;
      if (st == -1) {
        // the following point has no row: release it and go on without overlap (the tile just
        // finished has released the points it ended; at most the stream head is still held)
        st = build(false, nx);
        if (st == 1) {
          rn = embed_tile(nx, (tile + 1) & 1);
          hand_over();
        }
      }
      cur = nx; rcur = rn; ++tile; have = st == 1;
    }
    // tell the issuer warp that this group has no more tiles
    if (t == 0) s_flag[g] = 1;
    mbar_arrive(bar_a + g);
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

template <int BF16, int CULL>
__global__ void __launch_bounds__(TC_THREADS, 1) lvis_tc3_kernel(const TcParams p) {
  tc3_body<NF_MLP_LVIS, BF16, CULL>(p);
}
template <int BF16>
__global__ void __launch_bounds__(TC_THREADS, 1) brdf_tc3_kernel(const TcParams p) {
  tc3_body<NF_MLP_BRDF, BF16, 1>(p);
}

// ------------------------------------------------------------------ bring-up test
// One CTA: D[128x128] = A[128xK] * B[128xK]^T with A written to TMEM by tcgen05.st
// (the layout the epilogue uses) and B in the swizzle-free K-major image.
// a_host/b_host are row-major fp32 [128][K]; out is D row-major [128][128].
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, int K,
                     int swap_lbo_sbo, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint16_t* img = reinterpret_cast<uint16_t*>(smem);            // [K/8][128][8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)K * 128 * 2);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  for (int i = t; i < K * 128; i += 128) {
    int n = i / K, k = i % K;
    __half h = __float2half_rn(b[n * K + k]);
    img[((size_t)(k / 8) * 128 + n) * 8 + (k % 8)] = *reinterpret_cast<uint16_t*>(&h);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (t == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
  // A row t -> TMEM lane t, columns 128.. (16-bit pairs)
  for (int c0 = 0; c0 < K / 2; c0 += 8) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      pk[i] = pack2<0, 0>(a[t * K + 2 * (c0 + i)], a[t * K + 2 * (c0 + i) + 1]);
    TC_ST8(tb + 128 + c0, pk);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (t == 0) {
    uint32_t lbo = 128 * 16, sbo = 128;
    if (swap_lbo_sbo) { uint32_t x = lbo; lbo = sbo; sbo = x; }
    const uint32_t idesc = make_idesc(0, 128);
    for (int k = 0; k < K / 16; ++k)
      tc_mma_ts(tmem_base, tmem_base + 128 + k * 8,
                make_b_desc(smem_u32(img) + k * 2 * 128 * 16, lbo, sbo), idesc, k > 0);
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t r[32];
    TC_LD32(r, tb + cc * 32);
    tc_wait_ld();
    for (int i = 0; i < 32; ++i) out[t * 128 + cc * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}


// ---- bring-up test of the CTA-pair MMA: D[256 x 128] = A[256 x K] * B[128 x K]^T ----------
// cluster of 2 CTAs; CTA r holds rows [128 r, 128 r + 128) of A in its TMEM and rows
// [64 r, 64 r + 64) of B in its shared memory ([K/8][64][8] K-major, no swizzle).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, int K,
                      float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint16_t* img = reinterpret_cast<uint16_t*>(smem);            // [K/8][64][8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)K * 64 * 2);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, t = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  for (int i = t; i < K * 64; i += 128) {
    int n = i / K, k = i % K;
    __half h = __float2half_rn(b[(rank * 64 + n) * K + k]);
    img[((size_t)(k / 8) * 64 + n) * 8 + (k % 8)] = *reinterpret_cast<uint16_t*>(&h);
  }
  fence_proxy_async();
  if (t == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tc2_alloc(s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
  const float* arow = a + (size_t)(rank * 128 + t) * K;
  for (int c0 = 0; c0 < K / 2; c0 += 8) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack2<0, 0>(arow[2 * (c0 + i)], arow[2 * (c0 + i) + 1]);
    TC_ST8(tb + 128 + c0, pk);
  }
  tc_wait_st();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && t == 0) {
    const uint32_t idesc = make_idesc_mn(0, 256, 128);
    for (int k = 0; k < K / 16; ++k)
      tc2_mma_ts(tmem_base, tmem_base + 128 + k * 8,
                 make_b_desc(smem_u32(img) + k * 2 * 64 * 16, 64 * 16, 128), idesc, k > 0);
    tc2_commit_mc(bar, 3);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t r[32];
    TC_LD32(r, tb + cc * 32);
    tc_wait_ld();
    for (int i = 0; i < 32; ++i) out[(size_t)(rank * 128 + t) * 128 + cc * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { __syncwarp(); tc2_dealloc(tmem_base, 256); }
}

// host float -> fp16 / bf16 bits (round to nearest even)
uint16_t f2h_bits(float f) {
  __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
uint16_t f2bf_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}

template <int KIND, int BF16>
int launch_tc(nf_ctx* ctx, const nf_mlp* m, const TcParams& p, cudaStream_t st) {
  int grid = ctx->sm_count;
  int need = (p.n + 1) / 2;
  if (grid > need) grid = need;
  // NF_LVIS_V1=1 selects the first-generation kernel (bias added in the epilogue): A / B timing
  static const bool v1 = [] { const char* e = getenv("NF_LVIS_V1"); return e && e[0] == '1'; }();
  if (v1 && p.f_rgb == nullptr && p.xyz_dir == nullptr) {
    using SL = SmemLayout<KIND>;
    NF_CHECK_ARG(ctx, SL::total <= ctx->smem_optin, "shared memory budget exceeded");
    NF_CUDA(ctx, cudaFuncSetAttribute(mlp_tc_kernel<KIND, BF16>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::total));
    mlp_tc_kernel<KIND, BF16><<<grid, TC_THREADS, SL::total, st>>>(p);
  } else {
    // NF_LVIS_V2=1: the version-2 kernel for the visibility network too (A / B timing)
    static const bool v2 = [] { const char* e = getenv("NF_LVIS_V2"); return e && e[0] == '1'; }();
    // the learned-BRDF network runs on the version-3 machinery too (27.0 -> 20.6 ms at 640 k x 512,
    // bit-identical); NF_BRDF_V3=0 selects the version-2 kernel (group-issued MMAs) for A / B timing
    static const bool b3 = [] { const char* e = getenv("NF_BRDF_V3"); return !(e && e[0] == '0'); }();
    if (KIND == NF_MLP_BRDF && b3) {
      using S3 = SmemLayout3<NF_MLP_BRDF>;
      NF_CHECK_ARG(ctx, S3::total <= ctx->smem_optin, "shared memory budget exceeded");
      NF_CUDA(ctx, cudaFuncSetAttribute(brdf_tc3_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)S3::total));
      brdf_tc3_kernel<BF16><<<grid, TC_THREADS, S3::total, st>>>(p);
      NF_LAUNCH_CHECK(ctx);
      return NF_OK;
    }
    if (KIND == NF_MLP_LVIS && p.f_rgb == nullptr && !v2) {
      using S3 = SmemLayout3<NF_MLP_LVIS>;
      NF_CHECK_ARG(ctx, S3::total <= ctx->smem_optin, "shared memory budget exceeded");
      if (p.cull_normal) {
        NF_CUDA(ctx, cudaFuncSetAttribute(lvis_tc3_kernel<BF16, 1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S3::total));
        lvis_tc3_kernel<BF16, 1><<<grid, TC_THREADS, S3::total, st>>>(p);
      } else {
        NF_CUDA(ctx, cudaFuncSetAttribute(lvis_tc3_kernel<BF16, 0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S3::total));
        lvis_tc3_kernel<BF16, 0><<<grid, TC_THREADS, S3::total, st>>>(p);
      }
      NF_LAUNCH_CHECK(ctx);
      return NF_OK;
    }
    using SL = SmemLayout2<KIND>;
    NF_CHECK_ARG(ctx, SL::total <= ctx->smem_optin, "shared memory budget exceeded");
    constexpr int SELF = KIND == NF_MLP_BRDF ? 1 : 0;
    if (KIND == NF_MLP_BRDF || p.cull_normal) {
      NF_CUDA(ctx, cudaFuncSetAttribute(mlp_tc2_kernel<KIND, BF16, SELF, 1>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::total));
      mlp_tc2_kernel<KIND, BF16, SELF, 1><<<grid, TC_THREADS, SL::total, st>>>(p);
    } else {
      constexpr int C0 = KIND == NF_MLP_BRDF ? 1 : 0;     // (never taken for the BRDF network)
      NF_CUDA(ctx, cudaFuncSetAttribute(mlp_tc2_kernel<KIND, BF16, SELF, C0>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::total));
      mlp_tc2_kernel<KIND, BF16, SELF, C0><<<grid, TC_THREADS, SL::total, st>>>(p);
    }
  }
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// Host packing.  For the per-(point, light) networks (width 128, depth 4, skip
// after layer 2) append to the blob:
//   image (fp16) ++ image (bf16): segments [W0e | W1 | W2 | W3h | W3e], each
//     stored K-major without swizzle as [K/8][N=128][8] 16-bit values, i.e. the
//     element (n, k) of B = W^T sits at ((k/8)*128 + n)*16 + (k%8)*2 bytes.
//     W0e / W3e are the rows of W0 / W3 that multiply the per-row embedding
//     (light direction or Rusinkiewicz encoding), zero-padded to KE rows.
//   aux (fp32): b0..b3 [4][128], w_out [128], b_out [4], Wx0 [NR_PAD][128],
//     Wx3 [NR_PAD][128] = the rows of W0 / W3 that multiply the per-point columns.
int nf_sigma_tc_pack(nf_mlp* m);  // nf_sigma_tc.cu
int nf_point_tc_pack(nf_mlp* m);  // nf_point_tc.cu

int nf_tc_pack(nf_mlp* m) {
  const nf_mlp_desc& d = m->d;
  if (d.kind == NF_MLP_SIGMA) return nf_sigma_tc_pack(m);
  if (d.kind == NF_MLP_POINT) return nf_point_tc_pack(m);
  const bool pair_kind = d.kind == NF_MLP_LVIS || d.kind == NF_MLP_BRDF;
  if (!pair_kind || d.width != 128 || d.depth != 4 || d.skip_at != 2 || d.out_dim != 1) return NF_OK;
  const int KE = d.kind == NF_MLP_LVIS ? 32 : 16;
  const int NR_PAD = d.kind == NF_MLP_LVIS ? 64 : 8;
  const int nr = d.kind == NF_MLP_LVIS ? 3 * (1 + 2 * d.n_freqs_a) : d.z_dim;
  const int ne = d.in_dim - nr;  // per-row columns
  if (ne > KE || nr > NR_PAD) return NF_OK;  // unusual spec: FP32 path only
  if (d.kind == NF_MLP_LVIS && d.n_freqs_b != 4) return NF_OK;
  if (d.kind == NF_MLP_BRDF && d.n_freqs_a != 2) return NF_OK;
  const float* W[5];
  const float* B[5];
  for (int l = 0; l <= 4; ++l) { W[l] = d.W[l]; B[l] = d.b[l]; }
  const size_t halves = (size_t)(KE + 384 + KE + 32) * 128;   // + bias blocks of layers 1, 2
  const size_t aux_floats = AUX_WX0 + 2 * (size_t)NR_PAD * 128;
  size_t base = (m->blob.size() + 255) / 256 * 256;
  m->off_tc_f16 = base;
  m->off_tc_bf16 = base + (halves * 2 + 255) / 256 * 256;
  m->off_tc_aux = m->off_tc_bf16 + (halves * 2 + 255) / 256 * 256;
  m->tc_bytes = halves * 2;
  m->tc_aux_bytes = aux_floats * 4;
  m->blob.resize(m->off_tc_aux + (m->tc_aux_bytes + 255) / 256 * 256, 0);
  uint16_t* img16 = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tc_f16);
  uint16_t* imgbf = reinterpret_cast<uint16_t*>(m->blob.data() + m->off_tc_bf16);
  float* aux = reinterpret_cast<float*>(m->blob.data() + m->off_tc_aux);
  // segment s: rows of Keras W[l] (layout [K_l][128]) starting at row r0, count kc, padded to kp
  struct Seg { int l, r0, kc, kp; };
  const Seg segs[5] = {{0, nr, ne, KE}, {1, 0, 128, 128}, {2, 0, 128, 128}, {3, 0, 128, 128},
                       {3, 128 + nr, ne, KE}};
  size_t seg_base = 0;
  for (int s = 0; s < 5; ++s) {
    const Seg& sg = segs[s];
    for (int k = 0; k < sg.kp; ++k)
      for (int n = 0; n < 128; ++n) {
        float v = k < sg.kc ? W[sg.l][(size_t)(sg.r0 + k) * 128 + n] : 0.f;
        size_t idx = seg_base + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8);
        img16[idx] = f2h_bits(v);
        imgbf[idx] = f2bf_bits(v);
      }
    seg_base += (size_t)sg.kp * 128;
  }
  // bias blocks (version 2): a [16 k][128 n] K-major block per static-bias layer whose rows
  // k = 0 / 1 hold the bias as a 16-bit hi / lo pair (the A operand is the constant (1, 1, 0..))
  for (int l = 1; l <= 2; ++l) {
    for (int n = 0; n < 128; ++n) {
      const float bv = B[l][n];
      const uint16_t hi16 = f2h_bits(bv), hibf = f2bf_bits(bv);
      __half hh; memcpy(&hh, &hi16, 2);
      __nv_bfloat16 hb; memcpy(&hb, &hibf, 2);
      const size_t idx = seg_base + (size_t)n * 8;      // k-group 0, row n, k = 0
      // (hi, lo) at k = 0, 1 -- and repeated at k = 2..7 for the version-3 visibility kernel, whose
      // "ones" operand has its (1, 1) at k = 2 j, 2 j + 1 for a row of point j mod 4 (the other
      // kernels' ones operand is (1, 1, 0, ...): the copies meet zeros there)
      for (int j = 0; j < 4; ++j) {
        img16[idx + 2 * j] = hi16;
        img16[idx + 2 * j + 1] = f2h_bits(bv - __half2float(hh));
        imgbf[idx + 2 * j] = hibf;
        imgbf[idx + 2 * j + 1] = f2bf_bits(bv - __bfloat162float(hb));
      }
    }
    seg_base += (size_t)16 * 128;
  }
  for (int l = 0; l < 4; ++l) memcpy(aux + AUX_B + l * 128, B[l], 128 * sizeof(float));
  for (int c = 0; c < 128; ++c) aux[AUX_WOUT + c] = W[4][c];  // [128][1]
  aux[AUX_BOUT] = B[4][0];
  for (int k = 0; k < nr; ++k) {
    memcpy(aux + AUX_WX0 + (size_t)k * 128, W[0] + (size_t)k * 128, 128 * sizeof(float));
    memcpy(aux + AUX_WX0 + (size_t)(NR_PAD + k) * 128, W[3] + (size_t)(128 + k) * 128,
           128 * sizeof(float));
  }
  return NF_OK;
}

static int tc_common(nf_ctx* ctx, const nf_mlp* m, int precision, TcParams& p) {
  NF_CHECK_ARG(ctx, m->dev, "network not uploaded (call nf_mlp_upload first)");
  NF_CHECK_ARG(ctx, precision == NF_PREC_F16 || precision == NF_PREC_BF16, "bad precision");
  if (m->tc_bytes == 0)
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "no tcgen05 kernel for this network shape (need width 128, depth 4, "
                        "skip_at 2, out_dim 1); use NF_PREC_FP32");
  memset(&p, 0, sizeof(p));
  p.blob = (const uint8_t*)m->dev;
  p.off_img = precision == NF_PREC_BF16 ? m->off_tc_bf16 : m->off_tc_f16;
  p.off_aux = m->off_tc_aux;
  p.kind = m->d.kind;
  p.n_freqs_a = m->d.n_freqs_a; p.n_freqs_b = m->d.n_freqs_b; p.z_dim = m->d.z_dim;
  p.out_act = m->d.out_act;
  return NF_OK;
}

int nf_tc_lvis_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, int n, float xyz_scale,
                      const float* lxyz, int L, float* lvis, int precision, cudaStream_t st,
                      const float* xyz_dir, const float* cull_normal) {
  TcParams p;
  int rc = tc_common(ctx, m, precision, p);
  if (rc != NF_OK) return rc;
  NF_CHECK_ARG(ctx, L <= 1024, "n_lights > 1024 not supported by the tcgen05 kernel");
  if (n == 0) return NF_OK;
  p.n = n; p.L = L; p.nr = 3 * (1 + 2 * m->d.n_freqs_a); p.xyz_scale = xyz_scale;
  p.xyz = xyz; p.lxyz = lxyz; p.out = lvis; p.xyz_dir = xyz_dir; p.cull_normal = cull_normal;
  return precision == NF_PREC_BF16 ? launch_tc<NF_MLP_LVIS, 1>(ctx, m, p, st)
                                   : launch_tc<NF_MLP_LVIS, 0>(ctx, m, p, st);
}

// Light-visibility network with the rendering equation fused into the head epilogue
// (nf_stageB_fused_fwd): rgb[n,3]; lvis may be NULL.  Microfacet lobe, one env-map, L <= 512.
int nf_tc_lvis_render_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, int n, float xyz_scale,
                             const float* lxyz, int L, const float* normal, const float* cam,
                             const float* albedo, const float* rough, const float* lareas,
                             const float* light, const int* light_idx, float f0, int srgb,
                             float* lvis, float* rgb, int precision, cudaStream_t st) {
  TcParams p;
  int rc = tc_common(ctx, m, precision, p);
  if (rc != NF_OK) return rc;
  NF_CHECK_ARG(ctx, L <= 512, "fused rendering needs n_lights <= 512");
  if (n == 0) return NF_OK;
  p.n = n; p.L = L; p.nr = 3 * (1 + 2 * m->d.n_freqs_a); p.xyz_scale = xyz_scale;
  p.xyz = xyz; p.lxyz = lxyz; p.out = lvis;
  p.f_normal = normal; p.f_cam = cam; p.f_albedo = albedo; p.f_rough = rough; p.f_lareas = lareas;
  p.f_light = light; p.f_light_idx = light_idx; p.f_f0 = f0; p.f_srgb = srgb; p.f_rgb = rgb;
  return precision == NF_PREC_BF16 ? launch_tc<NF_MLP_LVIS, 1>(ctx, m, p, st)
                                   : launch_tc<NF_MLP_LVIS, 0>(ctx, m, p, st);
}

int nf_tc_brdf_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, const float* normal,
                      const float* cam, const float* z, int n, const float* lxyz, int L,
                      float* spec, int precision, cudaStream_t st) {
  TcParams p;
  int rc = tc_common(ctx, m, precision, p);
  if (rc != NF_OK) return rc;
  NF_CHECK_ARG(ctx, L <= 1024, "n_lights > 1024 not supported by the tcgen05 kernel");
  if (n == 0) return NF_OK;
  p.n = n; p.L = L; p.nr = m->d.z_dim; p.xyz_scale = 1.f;
  p.xyz = xyz; p.lxyz = lxyz; p.normal = normal; p.cam = cam; p.zlat = z; p.out = spec;
  return precision == NF_PREC_BF16 ? launch_tc<NF_MLP_BRDF, 1>(ctx, m, p, st)
                                   : launch_tc<NF_MLP_BRDF, 0>(ctx, m, p, st);
}

// Diagnostics (not part of the reference surface): single-tile tcgen05 self-test.
extern "C" int nf_selftest_umma(nf_ctx* ctx, const float* a_d, const float* b_d, int K,
                                int swap_lbo_sbo, float* out_d, void* stream) {
  NF_CHECK_ARG(ctx, a_d && b_d && out_d && K >= 16 && K <= 128 && K % 16 == 0, "bad argument");
  size_t sm = (size_t)K * 128 * 2 + 64;
  NF_CUDA(ctx, cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  umma_selftest_kernel<<<1, 128, sm, (cudaStream_t)stream>>>(a_d, b_d, K, swap_lbo_sbo, out_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

extern "C" int nf_selftest_umma2(nf_ctx* ctx, const float* a_d, const float* b_d, int K,
                                 float* out_d, void* stream) {
  NF_CHECK_ARG(ctx, a_d && b_d && out_d && K >= 16 && K <= 128 && K % 16 == 0, "bad argument");
  size_t sm = (size_t)K * 64 * 2 + 64;
  NF_CUDA(ctx, cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  umma2_selftest_kernel<<<2, 128, sm, (cudaStream_t)stream>>>(a_d, b_d, K, out_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
